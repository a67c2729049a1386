"""Developer tool / evidence for BASELINE config 4: the optimised chi-squared DAG (examples/chi_sq/src/main.rs:59-88, every
Multiply followed by its Relinearize) over B independent evaluations at n = 16384 — (a) node by node through the per-handle
FFI calls, as sunscreen_runtime issues them today, (b) DAG level by DAG level through the batch seams
(B200_Evaluator_*Batch), (c) the reference on one host thread.  Results of (a) and (b) are compared word for word."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from params import PARAMS
from sealc_driver import Sealc
import sealc_checks as sc
from sunscreen_b200.lib import B200Lib
vp, u64 = C.c_void_p, C.c_uint64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, moduli, t = PARAMS[os.environ.get("B200_PROBE", "n16384")]
S = Sealc(B200Lib.default().lib)
O = S.context(n, moduli, t)
kg, sk, pk, rlk, enc = vp(), vp(), vp(), vp(), vp()
S.call("KeyGenerator_Create1", O.ctx, C.byref(kg))
S.call("KeyGenerator_SecretKey", kg, C.byref(sk))
S.call("KeyGenerator_CreatePublicKey", kg, C.c_bool(False), C.byref(pk))
S.call("KeyGenerator_CreateRelinKeys", kg, C.c_bool(False), C.byref(rlk))
S.call("Encryptor_Create", O.ctx, pk, None, C.byref(enc))
rng = np.random.default_rng(3)
vals = rng.integers(1, 12, size=(B, 3))
def encrypt(v):
    h = O._dst(); S.call("Encryptor_Encrypt", enc, O.new_pt(np.array([v], dtype=np.uint64)), h, None); return h
t0 = time.perf_counter()
N0, N1, N2 = ([encrypt(int(vals[i, j])) for i in range(B)] for j in range(3))
print(f"encrypted {3*B} inputs in {time.perf_counter()-t0:.2f} s")
arr = lambda hs: (vp * len(hs))(*hs)
fresh = lambda k: [O._dst() for _ in range(k)]
def per_handle(i):
    mul = lambda a, b: O.relinearize(O.multiply(a, b), rlk)
    n0, n1, n2 = N0[i], N1[i], N2[i]
    x = O.add(O.add(n0, n0), n1); y = O.add(O.add(n2, n2), n1)
    a = mul(n0, n2); a = O.add(a, a); a = O.add(a, a)
    alpha = O.sub(a, mul(n1, n1)); alpha = mul(alpha, alpha)
    b1 = mul(x, x); b1 = O.add(b1, b1); b2 = mul(x, y); b3 = mul(y, y); b3 = O.add(b3, b3)
    return alpha, b1, b2, b3
def batched():
    def addsub(a, b, sub=False):
        d = fresh(len(a)); S.call("B200_Evaluator_AddSubBatch", O.ev, u64(len(a)), arr(a), arr(b), C.c_bool(sub), arr(d)); return d
    def mul(a, b):
        d = fresh(len(a)); S.call("B200_Evaluator_MultiplyRelinBatch", O.ev, u64(len(a)), arr(a), arr(b), rlk, arr(d)); return d
    t1 = addsub(N0 + N2, N0 + N2)                       # 2 n0 | 2 n2
    xy = addsub(t1, N1 + N1)                            # x | y
    X, Y = xy[:B], xy[B:]
    m = mul(N0 + N1 + X + X + Y, N2 + N1 + X + Y + Y)  # n0 n2 | n1^2 | x^2 | x y | y^2  : one launch sequence for 5B products
    a, n1sq, xx, xy_, yy = (m[i * B:(i + 1) * B] for i in range(5))
    d = addsub(a + xx + yy, a + xx + yy)                # 2a | b1 | b3
    a2, b1, b3 = d[:B], d[B:2 * B], d[2 * B:]
    a4 = addsub(a2, a2)
    alpha = addsub(a4, n1sq, True)
    alpha = mul(alpha, alpha)
    return alpha, b1, xy_, b3
words = lambda h: O.ct_words(h)
per_handle(0)
S.call("B200_SEALContext_Synchronize", O.ctx)
k = min(B, 32)
t0 = time.perf_counter(); ref_out = [per_handle(i) for i in range(k)]; dt_a = (time.perf_counter() - t0) / k
batched()
t0 = time.perf_counter(); out = batched(); dt_b = (time.perf_counter() - t0) / B
for i in (0, k - 1):
    for j in range(4):
        assert np.array_equal(words(ref_out[i][j]), words(out[j][i])), (i, j)
dec = vp(); S.call("Decryptor_Create", O.ctx, sk, C.byref(dec))
n0, n1, n2 = (int(v) for v in vals[0]); x, y = 2 * n0 + n1, 2 * n2 + n1
got = [int(O.pt_coeffs(O.decrypt(dec, out[j][0]))[0]) for j in range(4)]
assert got == [((4 * n0 * n2 - n1 * n1) ** 2) % t, (2 * x * x) % t, (x * y) % t, (2 * y * y) % t], got
print(f"chi_sq, n={n}, k={len(moduli)-1}: per-handle path {1/dt_a:8.1f} evaluations/s ({dt_a*1e3:.2f} ms each)")
print(f"chi_sq, n={n}, k={len(moduli)-1}: batch seams      {1/dt_b:8.1f} evaluations/s (B = {B}; outputs identical, decrypt correctly)")
try:
    import refseal
    R = refseal.RefContext(n, moduli, t)
    rkg = R.keygen(); rpk, rrk = R.public_key(rkg), R.relin_keys(rkg); renc = R.encryptor(rpk)
    rin = [R.encrypt(renc, R.new_pt(np.array([int(v)], dtype=np.uint64))) for v in vals[0]]
    def rcirc(n0, n1, n2):
        mul = lambda a, b: R.relinearize(R.multiply(a, b), rrk)
        x = R.add(R.add(n0, n0), n1); y = R.add(R.add(n2, n2), n1)
        a = mul(n0, n2); a = R.add(a, a); a = R.add(a, a)
        al = R.sub(a, mul(n1, n1)); al = mul(al, al)
        return al, R.add(mul(x, x), mul(x, x)), mul(x, y), mul(y, y)
    rcirc(*rin)
    t0 = time.perf_counter(); rcirc(*rin); rcirc(*rin); dt_r = (time.perf_counter() - t0) / 2
    print(f"chi_sq, reference, one host thread: {1/dt_r:8.2f} evaluations/s ({dt_r*1e3:.0f} ms each)")
except Exception as e:
    print("reference leg unavailable:", e)
