"""The Blake2xb generator behind key generation and encryption (sunscreen_b200/csrc/sampling.cpp) expands its output nodes
in SIMD lanes (AVX-512: 8, AVX2: 4) where the host has them.  Every width must produce the same byte stream as the scalar
restatement; the default width is compared with the reference's stream by the seeded-encryption parity tests
(tests/sealc_checks.py::encryption_components_parity)."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import ctypes as C, hashlib, sys
lib = C.CDLL(sys.argv[1])
f = lib._ZN4b2008blake2xbEPvmPKvmS2_m   # b200::blake2xb(void*, size_t, const void*, size_t, const void*, size_t)
f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
f.restype = None
h = hashlib.sha256()
for length in (1, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1000, 4096, 4097, 10000):
    for ctr in range(4):
        out = (C.c_uint8 * length)()
        key = (C.c_uint64 * 8)(*[(ctr * 77 + i * 1234567 + length) & (2**64 - 1) for i in range(8)])
        msg = C.c_uint64(ctr)
        f(out, length, C.byref(msg), 8, key, 64)
        h.update(bytes(out))
print(h.hexdigest())
"""


def test_simd_widths_agree(emu_lib):
    path = os.path.join(ROOT, "tests", "emu", "_build", "libb200bfv_emu.so")
    digests = {}
    for width in ("1", "4", "8"):
        env = dict(os.environ, B200_PRNG_SIMD=width)
        out = subprocess.run([sys.executable, "-c", _CHILD, path], env=env, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        digests[width] = out.stdout.strip()
    assert len(set(digests.values())) == 1, digests
    assert len(digests["1"]) == 64
