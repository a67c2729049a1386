"""Developer tool: latency of the SURVEY §8 row E entry points through the FFI — key generation, encryption, batch encoding —
ours vs the reference on one host thread.  Sampling stays on the host (bit-exact PRNG streams); the NTT / dyadic arithmetic
around it runs on the GPU."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sealc_checks as sc, refseal
from params import PARAMS
from sealc_driver import Sealc
from sunscreen_b200.lib import B200Lib
S = Sealc(B200Lib.default().lib)


def timed(f, reps):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps * 1e3


for name in sys.argv[1:] or ("n8192", "n16384"):
    n, moduli, t = PARAMS[name]
    R = refseal.RefContext(n, moduli, t); O = S.context(n, moduli, t)
    for L, who in zip(sc._libs(R, O), ("reference", "ours     ")):
        kg = C.c_void_p(); L.call("KeyGenerator_Create1", L.ctx, C.byref(kg))
        sk = C.c_void_p(); L.call("KeyGenerator_SecretKey", kg, C.byref(sk))
        keep = []

        def mk(fn):
            def f():
                h = C.c_void_p(); L.call(fn, kg, C.c_bool(False), C.byref(h)); keep.append(h)
            return f
        t_pk = timed(mk("KeyGenerator_CreatePublicKey"), 10)
        pk = keep[-1]
        t_rlk = timed(mk("KeyGenerator_CreateRelinKeys"), 5)
        t_glk = timed(mk("KeyGenerator_CreateGaloisKeysAll"), 1)
        enc = C.c_void_p(); L.call("Encryptor_Create", L.ctx, pk, sk, C.byref(enc))
        be = C.c_void_p(); L.call("BatchEncoder_Create", L.ctx, C.byref(be))
        vals = (np.arange(n, dtype=np.uint64) % t).astype(np.uint64)
        pt = L.new("Plaintext"); ct = L.new("Ciphertext")
        vp = vals.ctypes.data_as(C.c_void_p)
        t_be = timed(lambda: L.call("BatchEncoder_Encode1", be, C.c_uint64(n), vp, pt), 20)
        cnt = C.c_uint64(0); out = np.zeros(n, dtype=np.uint64)
        t_bd = timed(lambda: L.call("BatchEncoder_Decode1", be, pt, C.byref(cnt), out.ctypes.data_as(C.c_void_p), None), 20)
        assert np.array_equal(out, vals)
        t_enc = timed(lambda: L.call("Encryptor_Encrypt", enc, pt, ct, None), 20)
        t_sym = timed(lambda: L.call("Encryptor_EncryptSymmetric", enc, pt, C.c_bool(False), ct, None), 20)
        print(f"{name} {who}: public key {t_pk:7.2f}  relin keys {t_rlk:8.2f}  galois keys (all) {t_glk:9.1f}  encode {t_be:6.3f}  "
              f"decode {t_bd:6.3f}  encrypt {t_enc:6.2f}  encrypt symmetric {t_sym:6.2f}   ms", flush=True)
