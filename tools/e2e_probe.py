"""Developer tool: host-buffer (e2e) throughput of multiply_relin vs raw PCIe copy bandwidth on this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunscreen_b200.lib import B200Context
from bench import MODULI, PLAIN, N_POLY
B = 1024
ctx = B200Context(N_POLY, MODULI, PLAIN)
k = ctx.k(); n = N_POLY
ah = torch.randint(0, MODULI[0], (B, 2, k, n), dtype=torch.int64).pin_memory()
bh = torch.randint(0, MODULI[0], (B, 2, k, n), dtype=torch.int64).pin_memory()
oh = torch.empty((B, 2, k, n), dtype=torch.int64).pin_memory()
rlk = torch.randint(0, MODULI[0], (k, 2, 5, n), dtype=torch.int64, device="cuda")
d = torch.empty((B, 2, k, n), dtype=torch.int64, device="cuda")
for _ in range(2):
    d.copy_(ah, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(ah, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"H2D {ah.numel()*8/(t1-t0)/1e9:.1f} GB/s")
t0 = time.perf_counter(); oh.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"D2H {ah.numel()*8/(t1-t0)/1e9:.1f} GB/s")
for chunk in (32, 64, 128, 256):
    os.environ["B200_HOST_CHUNK"] = str(chunk)
    ctx.multiply_relin_host(ah, bh, rlk, oh, B)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.multiply_relin_host(ah, bh, rlk, oh, B)
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk {chunk}: {B/dt:.0f} ops/s ({dt*1e3:.1f} ms per {B}; {3*B*2*k*n*8/dt/1e9:.1f} GB/s moved)")
