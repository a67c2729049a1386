"""Developer tool: run only the batched NTT (BASELINE config 2) a few times — used under ncu."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunscreen_b200.lib import B200Context
from bench import MODULI, PLAIN, N_POLY
items = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = B200Context(N_POLY, MODULI, PLAIN)
k = ctx.k()
x = torch.empty((items, k, N_POLY), dtype=torch.int64, device="cuda")
for i in range(k):
    x[:, i, :] = torch.randint(0, MODULI[i], (items, N_POLY), device="cuda", dtype=torch.int64)
s = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    ctx.ntt_forward(x, items, stream=s)
    ctx.ntt_inverse(x, items, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.ntt_forward(x, items, stream=s); e1.record(); torch.cuda.synchronize()
f = e0.elapsed_time(e1)
e0.record(); ctx.ntt_inverse(x, items, stream=s); e1.record(); torch.cuda.synchronize()
i_ = e0.elapsed_time(e1)
b = 16 * N_POLY * items * k
print(f"fwd {f:.3f} ms {b/f/1e6:.0f} GB/s | inv {i_:.3f} ms {b/i_/1e6:.0f} GB/s")
