"""Developer tool / evidence for BASELINE config 5: rotate_rows followed by multiply_plain at n = 32768 (k = 15 + special),
device-resident batch through the layer-1 ABI, next to the reference on one host thread."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from params import PARAMS
from sunscreen_b200.lib import B200Context
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, moduli, t = PARAMS[os.environ.get("B200_PROBE", "n32768")]
ctx = B200Context(n, moduli, t)
k = ctx.k()
def rr(shape, mods):
    out = torch.empty(shape + (len(mods), n), dtype=torch.int64, device="cuda")
    for i, m in enumerate(mods):
        out[..., i, :] = torch.randint(0, m, shape + (n,), device="cuda", dtype=torch.int64)
    return out
a = rr((B, 2), moduli[:k]); key = rr((k, 2), moduli)
plain = torch.randint(0, t, (B, n), device="cuda", dtype=torch.int64)
o1, o2 = torch.zeros_like(a), torch.zeros_like(a)
elt = ctx.galois_elt_from_step(1)
s = torch.cuda.current_stream().cuda_stream
def step():
    ctx.apply_galois(a, elt, key, o1, B, stream=s)
    ctx.multiply_plain(o1, 2, plain, B, o2, B, stream=s)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"n={n} k={k}: rotate_rows + multiply_plain, batch {B}: {ms:.2f} ms/step, {B/ms*1e3:.0f} pairs/s")
try:
    import refseal
    R = refseal.RefContext(n, moduli, t)
    kg = R.keygen(); pk = R.public_key(kg); glk = R.galois_keys_steps(kg, [1]); enc = R.encryptor(pk)
    ct = R.encrypt(enc, R.new_pt(np.arange(1, 50, dtype=np.uint64)))
    pl = R.new_pt(np.random.default_rng(1).integers(1, t, size=n, dtype=np.uint64))
    R.multiply_plain(R.rotate_rows(ct, 1, glk), pl)
    t0 = time.perf_counter()
    for _ in range(2): R.multiply_plain(R.rotate_rows(ct, 1, glk), pl)
    dt = (time.perf_counter() - t0) / 2
    print(f"reference, one host thread: {1/dt:.2f} pairs/s ({dt*1e3:.0f} ms each)")
except Exception as e:
    print("reference leg unavailable:", e)
