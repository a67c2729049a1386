"""Developer tool: latency of Decryptor_InvariantNoiseBudget / Decryptor_Decrypt through the FFI (sunscreen_runtime calls the
budget before every decrypt, runtime.rs:182), ours vs the reference on one host thread."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sealc_checks as sc, refseal
from params import PARAMS
from sealc_driver import Sealc
from sunscreen_b200.lib import B200Lib
S = Sealc(B200Lib.default().lib)
for name in ("n8192", "n16384"):
    n, moduli, t = PARAMS[name]
    R = refseal.RefContext(n, moduli, t); O = S.context(n, moduli, t)
    RL, OL = sc._libs(R, O)
    kg = R.keygen(); sk, pk = R.secret_key(kg), R.public_key(kg)
    ct = R.encrypt(R.encryptor(pk), R.new_pt(np.arange(1, 9, dtype=np.uint64)))
    oct_ = OL.load("Ciphertext", RL.save("Ciphertext", ct, 0))
    osk = OL.load("SecretKey", RL.save("SecretKey", sk, 0))
    rdec = R.decryptor(sk); odec = C.c_void_p(); O.S.call("Decryptor_Create", O.ctx, osk, C.byref(odec))
    for L, d, h, who in ((RL, rdec, ct, "reference"), (OL, odec, oct_, "ours     ")):
        b = C.c_int(); out = L.new("Plaintext")
        L.call("Decryptor_InvariantNoiseBudget", d, h, C.byref(b)); L.call("Decryptor_Decrypt", d, h, out)
        t0 = time.perf_counter()
        for _ in range(20): L.call("Decryptor_InvariantNoiseBudget", d, h, C.byref(b))
        t1 = time.perf_counter()
        for _ in range(20): L.call("Decryptor_Decrypt", d, h, out)
        t2 = time.perf_counter()
        print(f"{name} {who}: noise budget {(t1-t0)/20*1e3:6.2f} ms   decrypt {(t2-t1)/20*1e3:6.2f} ms   (budget {b.value})")
