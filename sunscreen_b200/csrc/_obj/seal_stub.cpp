extern "C" const char b200_libseal_stub[] = "libseal-4.0.a: the B200 backend lives in libsealc-4.0.a";
