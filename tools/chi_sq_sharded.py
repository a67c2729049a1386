"""BASELINE config 4: the optimised chi-squared DAG (examples/chi_sq/src/main.rs:59-88; every Multiply followed by the
Relinearize that insert_relinearizations.rs:17-62 emits) at n = 16384, k = 8, over B = 256 independent evaluations,
SHARDED over the GPUs of one node the way the north star asks: rank 0 holds the batch (fresh encryptions), NCCL scatter of
the inputs -> every rank evaluates its slice DAG level by DAG level through the plugin's batch seams
(B200_Evaluator_{AddSub,MultiplyRelin}Batch) -> NCCL gather of the 4 outputs to rank 0.  Keys are replicated (serialised
on rank 0, loaded by every rank).  Checked: gathered outputs == the same evaluations done on ONE GPU, word for word, and the
first evaluation decrypts to the circuit's values.

  python tools/chi_sq_sharded.py [B]                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node N ... tools/chi_sq_sharded.py [B]
Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from params import PARAMS
from sealc_driver import Sealc
from sunscreen_b200.lib import B200Lib

vp, u64 = C.c_void_p, C.c_uint64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
os.environ["B200_DEVICE"] = str(local)
n, moduli, t = PARAMS[os.environ.get("B200_PROBE", "n16384")]
k = len(moduli) - 1
S = Sealc(B200Lib.default().lib)
O = S.context(n, moduli, t)
arr = lambda hs: (vp * len(hs))(*hs)
fresh = lambda cnt: [O._dst() for _ in range(cnt)]
ptr = lambda tsr: C.cast(tsr.data_ptr(), C.POINTER(u64))
ct_words = 2 * k * n


def save(kind, h):
    size = C.c_int64()
    S.call(kind + "_SaveSize", h, C.c_uint8(0), C.byref(size))
    buf = (C.c_uint8 * size.value)()
    out = C.c_int64()
    S.call(kind + "_Save", h, buf, u64(size.value), C.c_uint8(0), C.byref(out))
    return bytes(buf[: out.value])


def bcast_bytes(b):
    if dist is None:
        return b
    ln = torch.tensor([len(b) if rank == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(ln, 0)
    buf = torch.empty(int(ln.item()), dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy())


# ---- keys on rank 0, replicated ----
rlk_bytes = sk_h = None
vals = None
if rank == 0:
    kg, sk_h, pk, rlk0, enc = vp(), vp(), vp(), vp(), vp()
    S.call("KeyGenerator_Create1", O.ctx, C.byref(kg))
    S.call("KeyGenerator_SecretKey", kg, C.byref(sk_h))
    S.call("KeyGenerator_CreatePublicKey", kg, C.c_bool(False), C.byref(pk))
    S.call("KeyGenerator_CreateRelinKeys", kg, C.c_bool(False), C.byref(rlk0))
    S.call("Encryptor_Create", O.ctx, pk, None, C.byref(enc))
    rlk_bytes = save("KSwitchKeys", rlk0)
rlk_bytes = bcast_bytes(rlk_bytes)
rlk = vp()
S.call("KSwitchKeys_Create1", C.byref(rlk))
inb = C.c_int64()
S.call("KSwitchKeys_Load", rlk, O.ctx, (C.c_uint8 * len(rlk_bytes)).from_buffer_copy(rlk_bytes), u64(len(rlk_bytes)), C.byref(inb))

# ---- the batch lives on rank 0: B x (n0, n1, n2) fresh encryptions as raw words on its GPU ----
full_in = None
if rank == 0:
    rng = np.random.default_rng(3)
    vals = rng.integers(1, 12, size=(B, 3))
    hs = []
    for i in range(B):
        for j in range(3):
            h = O._dst()
            S.call("Encryptor_Encrypt", enc, O.new_pt(np.array([int(vals[i, j])], dtype=np.uint64)), h, None)
            hs.append(h)
    full_in = torch.empty((B, 3, ct_words), dtype=torch.int64, device=dev)
    S.call("B200_Ciphertext_GetWordsBatch", O.ctx, u64(3 * B), arr(hs), ptr(full_in), u64(full_in.numel()))
per = B // world
assert per * world == B, "B must be a multiple of the number of GPUs"


def evaluate(inp, cnt):
    """inp: (cnt, 3, ct_words) device tensor -> (cnt, 4, ct_words) device tensor; one launch sequence per DAG level"""
    H = [fresh(cnt) for _ in range(3)]
    flat = inp.transpose(0, 1).contiguous()  # (3, cnt, words): operand-major so each operand is one SetWordsBatch
    torch.cuda.synchronize()  # the plugin works on its own streams: torch's copy must have landed before it reads `flat`
    for j in range(3):
        S.call("B200_Ciphertext_SetWordsBatch", O.ctx, u64(cnt), arr(H[j]), O.first_id, u64(2), C.c_bool(False), ptr(flat[j]))
    N0, N1, N2 = H

    def addsub(a, b, sub=False):
        d = fresh(len(a))
        S.call("B200_Evaluator_AddSubBatch", O.ev, u64(len(a)), arr(a), arr(b), C.c_bool(sub), arr(d))
        return d

    def mul(a, b):
        d = fresh(len(a))
        S.call("B200_Evaluator_MultiplyRelinBatch", O.ev, u64(len(a)), arr(a), arr(b), rlk, arr(d))
        return d

    t1 = addsub(N0 + N2, N0 + N2)
    xy = addsub(t1, N1 + N1)
    X, Y = xy[:cnt], xy[cnt:]
    m = mul(N0 + N1 + X + X + Y, N2 + N1 + X + Y + Y)  # n0 n2 | n1^2 | x^2 | x y | y^2
    a, n1sq, xx, xy_, yy = (m[i * cnt:(i + 1) * cnt] for i in range(5))
    d = addsub(a + xx + yy, a + xx + yy)
    a2, b1, b3 = d[:cnt], d[cnt:2 * cnt], d[2 * cnt:]
    a4 = addsub(a2, a2)
    al0 = addsub(a4, n1sq, True)
    alpha = mul(al0, al0)
    out = torch.empty((4, cnt, ct_words), dtype=torch.int64, device=dev)
    for j, hs_ in enumerate((alpha, b1, xy_, b3)):
        S.call("B200_Ciphertext_GetWordsBatch", O.ctx, u64(cnt), arr(hs_), ptr(out[j]), u64(out[j].numel()))
    for group in (N0, N1, N2, t1, xy, m, d, a4, al0, alpha):
        for h in group:
            S.call("Ciphertext_Destroy", h)
    return out.transpose(0, 1).contiguous()


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


loc_in = torch.empty((per, 3, ct_words), dtype=torch.int64, device=dev)
gathered = [torch.empty((per, 4, ct_words), dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None


def sharded_step():
    t0 = time.perf_counter()
    if dist is not None:
        dist.scatter(loc_in, list(full_in.chunk(world)) if rank == 0 else None, src=0)
        torch.cuda.synchronize()
    else:
        loc_in.copy_(full_in)
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    o = evaluate(loc_in, per)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if dist is not None:
        dist.gather(o, gathered, dst=0)
        torch.cuda.synchronize()
    else:
        gathered[0].copy_(o)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, t3 - t0


sharded_step()  # warm
barrier()
reps = 3
acc = np.zeros(4)
for _ in range(reps):
    barrier()
    acc += np.array(sharded_step())
acc /= reps
if dist is not None:
    tt = torch.tensor(acc, dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    acc = tt.cpu().numpy()
if rank == 0:
    whole = torch.cat(gathered)
    ok = None
    one_gpu = None
    if world > 1:  # the same evaluations on ONE GPU (rank 0 alone), in slices of `per`
        t0 = time.perf_counter()
        ref = torch.cat([evaluate(full_in[i:i + per], per) for i in range(0, B, per)])
        torch.cuda.synchronize()
        one_gpu = time.perf_counter() - t0
        ok = bool(torch.equal(ref, whole))
    # first evaluation decrypts to the circuit's values
    dec = vp()
    S.call("Decryptor_Create", O.ctx, sk_h, C.byref(dec))
    outs = fresh(4)
    first = whole[0].contiguous()
    S.call("B200_Ciphertext_SetWordsBatch", O.ctx, u64(4), arr(outs), O.first_id, u64(2), C.c_bool(False), ptr(first))
    got = [int(O.pt_coeffs(O.decrypt(dec, h))[0]) for h in outs]
    n0, n1, n2 = (int(v) for v in vals[0])
    x, y = 2 * n0 + n1, 2 * n2 + n1
    want = [((4 * n0 * n2 - n1 * n1) ** 2) % t, (2 * x * x) % t, (x * y) % t, (2 * y * y) % t]
    sc, comp, ga, tot = (float(v) for v in acc)
    print(json.dumps({
        "workload": f"chi_sq DAG (6 multiply+relinearize, 9 add/sub, 4 outputs), n={n}, k={k}, {B} evaluations held by rank 0",
        "n_gpus": world, "evaluations_per_s": B / tot, "evaluations_per_s_compute_only": B / comp, "scatter_ms": 1e3 * sc,
        "compute_ms": 1e3 * comp, "gather_ms": 1e3 * ga, "total_ms": 1e3 * tot,
        "scatter_bytes": (world - 1) * per * 3 * ct_words * 8, "gather_bytes": (world - 1) * per * 4 * ct_words * 8,
        "equals_one_gpu_words": ok, "one_gpu_same_slices_s": one_gpu, "decrypts_correctly": got == want,
        "timing": "wall clock between synchronisations, max over ranks, mean of %d steps" % reps}))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
