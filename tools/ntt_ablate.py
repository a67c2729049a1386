"""Developer tool: time the FP64 NTT kernel variants (NttFpStaticPass VAR, ntt_fp_body.cuh) on BASELINE config 2
(4096 x 4 residue polynomials, n = 8192), CUDA events, best of `reps`.

  VAR 0  the shipping kernel with twiddles from global memory
  VAR 1  first 512 twiddles of the table in shared memory            (results checked against VAR 0)
  VAR 16 streaming: persistent CTAs, next polynomial requested during the last pass (results checked against VAR 0)
  VAR 2 / 4 / 8 and sums: ABLATIONS — no twiddle loads / one-DMUL products / no global traffic.  Their results are
  meaningless; they measure what each component of the kernel costs (which is what bounds it).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunscreen_b200.lib import B200Context, B200Lib
from bench import MODULI, PLAIN, N_POLY

items = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else [0, 1, 16, 2, 4, 6, 8, 14]
lib = B200Lib.default()
ctx = B200Context(N_POLY, MODULI, PLAIN)
k = ctx.k()
x0 = torch.empty((items, k, N_POLY), dtype=torch.int64, device="cuda")
for i in range(k):
    x0[:, i, :] = torch.randint(0, MODULI[i], (items, N_POLY), device="cuda", dtype=torch.int64)
s = torch.cuda.current_stream().cuda_stream
bytes_ = 16 * N_POLY * items * k


def timed(fn):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ref_f = ref_i = None
# "16:6000" = variant 16 with a start-up stagger of 6000 clock cycles per resident CTA slot
for spec in variants:
    var, _, stag = str(spec).partition(":")
    var = int(var)
    lib.lib.b200_debug_ntt_stagger(int(stag or 0))
    lib.lib.b200_debug_ntt_variant(var)
    x = x0.clone()
    ctx.ntt_forward(x, items, stream=s)
    torch.cuda.synchronize()
    fwd_out = x.clone()
    ctx.ntt_inverse(x, items, stream=s)
    torch.cuda.synchronize()
    note = ""
    if var == 0:
        ref_f = fwd_out
        assert torch.equal(x, x0), "round trip"
    elif var in (1, 16, 2048, 2049):
        note = "  forward == VAR 0: %s, round trip: %s" % (torch.equal(fwd_out, ref_f), torch.equal(x, x0))
    for _ in range(2):
        ctx.ntt_forward(x, items, stream=s)
        ctx.ntt_inverse(x, items, stream=s)
    f = timed(lambda: ctx.ntt_forward(x, items, stream=s))
    i_ = timed(lambda: ctx.ntt_inverse(x, items, stream=s))
    print(f"VAR {str(spec):>8s}: fwd {f:.3f} ms {bytes_ / f / 1e6:6.0f} GB/s | inv {i_:.3f} ms {bytes_ / i_ / 1e6:6.0f} GB/s{note}", flush=True)
lib.lib.b200_debug_ntt_variant(-1)
lib.lib.b200_debug_ntt_stagger(0)
