"""world_size-2 gloo test of the multi-GPU host logic (scatter -> independent per-rank compute -> gather), with the
CPU emulation build standing in for the per-rank GPU library."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from sunscreen_b200.lib import B200Lib, B200Context
from sunscreen_b200.sharding import sharded_multiply_relin, shard_bounds
from params import PARAMS
import parity_checks as pc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = B200Lib(os.path.join(sys.argv[1], "tests", "emu", "_build", "libb200bfv_emu.so"), _allow_emu=True)
n, moduli, t = PARAMS["n4096"]
ctx = B200Context(n, moduli, t, lib=emu)
k = ctx.k()
rng = np.random.default_rng(5)          # same stream on every rank: rank 0's copy is "the" batch
B = 5
A = pc.rand_ct(rng, moduli, k, n, batch=B); Bc = pc.rand_ct(rng, moduli, k, n, batch=B)
key = pc.rand_ksk(rng, moduli, k, n)
ta = torch.from_numpy(A.view(np.int64)); tb = torch.from_numpy(Bc.view(np.int64)); tk = torch.from_numpy(key.view(np.int64))
out = sharded_multiply_relin(ctx, ta if rank == 0 else None, tb if rank == 0 else None, tk, B, "cpu")
assert [shard_bounds(B, r, 2) for r in range(2)] == [(0, 2), (2, 5)]
if rank == 0:
    ref = torch.zeros((B, 2, k, n), dtype=torch.int64)
    ctx.multiply_relin(ta, tb, tk, ref, B)
    assert torch.equal(out, ref), "sharded result differs from single-rank result"
    print("SHARD_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_scatter_compute_gather_world2(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script), ROOT], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK" in r.stdout
