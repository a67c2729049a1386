"""Developer tool: per-CTA phase timeline of the forward / inverse NTT kernels (b200_ntt_timeline)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sunscreen_b200.lib import B200Context
from bench import MODULI, PLAIN, N_POLY
items = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = B200Context(N_POLY, MODULI, PLAIN)
k = ctx.k()
x = torch.empty((items, k, N_POLY), dtype=torch.int64, device="cuda")
for i in range(k):
    x[:, i, :] = torch.randint(0, MODULI[i], (items, N_POLY), device="cuda", dtype=torch.int64)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.ntt_forward(x, items, stream=s); ctx.ntt_inverse(x, items, stream=s)
torch.cuda.synchronize()
ctas = items * k
tl = torch.zeros((ctas, 8), dtype=torch.int64, device="cuda")
fn = ctx.L.lib.b200_ntt_timeline
fn.argtypes = [C.c_void_p, C.c_void_p]; fn.restype = None
for name, op in (("fwd", ctx.ntt_forward), ("inv", ctx.ntt_inverse)):
    tl.zero_()
    fn(ctx.h, C.c_void_p(tl.data_ptr()))
    op(x, items, stream=s)
    torch.cuda.synchronize()
    fn(ctx.h, None)
    t = tl.cpu().numpy().astype(np.int64)
    t0 = t[:, 1].min()
    # columns: 0 smid, 1 start, 2..5 after pass 1..4, 6 copy-in done (staged input only, else 0), 7 end
    start, p1, p2, p3, p4, cin, end = (t[:, i] - t0 for i in (1, 2, 3, 4, 5, 6, 7))
    has_cin = t[:, 6].max() > 0
    span = end.max()
    print(f"== {name}: {ctas} CTAs, kernel span {span/1e3:.1f} us, {ctas/444:.1f} waves of 444")
    for label, sel in (("first wave", start < 1000), ("steady state", (start > 0.3 * span) & (start < 0.7 * span))):
        if not sel.any():
            continue
        m = lambda a: a[sel].mean() / 1e3
        parts = []
        if has_cin:
            parts.append(f"copy-in {m(cin - start):5.2f}")
            parts.append(f"pass1 {m(p1 - cin):5.2f}")
        else:
            parts.append(f"pass1(+loads) {m(p1 - start):5.2f}")
        parts += [f"pass2 {m(p2 - p1):5.2f}", f"pass3 {m(p3 - p2):5.2f}", f"pass4 {m(p4 - p3):5.2f}", f"tail/copy-out {m(end - p4):5.2f}",
                  f"total {m(end - start):5.2f}"]
        print(f"   {label:12s} ({int(sel.sum())} CTAs), us: " + "  ".join(parts))
