"""Developer tool: device-resident multiply_relin throughput for a few B200_MR_SPLIT settings (one process each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunscreen_b200.lib import B200Context
from bench import MODULI, PLAIN, N_POLY
if os.environ.get("B200_PROBE"):  # e.g. n8192_54, n16384, n32768 (tests/params.py)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from params import PARAMS
    N_POLY, MODULI, PLAIN = PARAMS[os.environ["B200_PROBE"]]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = B200Context(N_POLY, MODULI, PLAIN)
k = ctx.k(); n = N_POLY
def rr(shape, mods):
    out = torch.empty(shape + (len(mods), n), dtype=torch.int64, device="cuda")
    for i, m in enumerate(mods):
        out[..., i, :] = torch.randint(0, m, shape + (n,), device="cuda", dtype=torch.int64)
    return out
a, b, rlk = rr((B, 2), MODULI[:k]), rr((B, 2), MODULI[:k]), rr((k, 2), MODULI)
out = torch.zeros_like(a)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.multiply_relin(a, b, rlk, out, B, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.multiply_relin(a, b, rlk, out, B, stream=s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"split={os.environ.get('B200_MR_SPLIT','1')} batch={B}: {ms:.3f} ms/step {B/ms*1e3:.0f} ops/s  checksum {int(out.sum().item()) & 0xffffffff:08x}")
if os.environ.get("B200_TRACE"):
    ctx.L.lib.b200_trace_dump()  # per-kernel totals over the 3 warm-up + 5 timed steps
