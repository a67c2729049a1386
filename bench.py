#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: ciphertext multiply+relinearize per second at n=8192,
218-bit BFVDefault chain (k=4 data residues + special prime), batch = 1024 ciphertext pairs per GPU.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W  # the reference's own CPU implementation (oracle/_ref)

One "step" = one pass of the hot path (Evaluator::multiply then ::relinearize) over the whole batch.
`value` = device-resident throughput (inputs already in HBM), `e2e` = the same metric through the host-buffer
C-ABI entry point (pinned host memory in, host memory out, copies inside the timed region).
`roofline` is measured live on the dominant kernel (the batched NTT, BASELINE config 2: 4096 polynomials x 4
residues) with CUDA events on the launching stream; `cpu_baseline` times the unmodified reference on this box's
host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_POLY = 8192
MODULI = [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001]
PLAIN = 1032193
BATCH = 1024
NTT_POLYS = 4096  # BASELINE config 2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.samples = []
        self.mark = 0
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for s in self.samples[max(self.mark - 1, 0):]:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so pinned host buffers are node-local."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = "0000:" + bdf[9:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_reference_rate(threads, iters, warmup=2):
    """Unmodified reference (oracle/_ref/libsealc_ref.so): threads x iters multiply+relinearize_inplace, own pool per thread."""
    import numpy as np
    import refseal
    R = refseal.RefContext(N_POLY, MODULI, PLAIN)
    rng = np.random.default_rng(0)
    k = 4

    def rand_ct():
        w = np.empty((2, k, N_POLY), dtype=np.uint64)
        for i in range(k):
            w[:, i, :] = rng.integers(0, MODULI[i], size=(2, N_POLY), dtype=np.uint64)
        return R.new_ct(w)

    key = np.empty((k, 2, 5, N_POLY), dtype=np.uint64)
    for i in range(5):
        key[:, :, i, :] = rng.integers(0, MODULI[i], size=(k, 2, N_POLY), dtype=np.uint64)
    a, b, rlk = rand_ct(), rand_ct(), R.new_ksk({0: key})
    secs = R.time_mul_relin(a, b, rlk, threads, iters, warmup)
    return threads * iters / secs, secs


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref) on all host cores; one step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import refseal
    cores = os.cpu_count() or 1
    iters = 16  # per thread per step: cores*16 multiply+relinearize per step
    R = refseal.RefContext(N_POLY, MODULI, PLAIN)
    rng = np.random.default_rng(0)
    k = 4
    w = np.empty((2, 2, k, N_POLY), dtype=np.uint64)
    key = np.empty((k, 2, 5, N_POLY), dtype=np.uint64)
    for i in range(k):
        w[:, :, i, :] = rng.integers(0, MODULI[i], size=(2, 2, N_POLY), dtype=np.uint64)
    for i in range(5):
        key[:, :, i, :] = rng.integers(0, MODULI[i], size=(k, 2, N_POLY), dtype=np.uint64)
    a, b, rlk = R.new_ct(w[0]), R.new_ct(w[1]), R.new_ksk({0: key})
    for _ in range(max(args.warmup, 1)):
        R.time_mul_relin(a, b, rlk, cores, 4, 1)
    secs = 0.0
    for _ in range(args.steps):
        secs += R.time_mul_relin(a, b, rlk, cores, iters, 0)
    rate = args.steps * cores * iters / secs
    line = {
        "metric": "ciphertext mul+relin/sec", "value": rate, "unit": "ops/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "Evaluator::multiply + relinearize, n=8192, BFVDefault 218-bit chain (k=4 + special), "
                               f"bounded sample of {cores * iters} ct pairs per step", "batch_per_step": cores * iters,
                   "parallelism": f"{cores} host threads, one memory pool each (oracle/_ref = unmodified reference)"},
        "cpu_baseline": {"value": rate, "unit": "ops/s", "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} steps x {cores} threads x {iters} multiply+relinearize_inplace, "
                                   "uniform-random ciphertext words"},
        "e2e": {"value": rate, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from sunscreen_b200.lib import B200Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None if os.environ.get("B200_BENCH_NO_NUMA") else bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    ctx = B200Context(N_POLY, MODULI, PLAIN, device=local)
    k = ctx.k()
    n = N_POLY
    B = args.batch
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)

    def rand_rows(shape_prefix, mods):
        out = torch.empty(shape_prefix + (len(mods), n), dtype=torch.int64, device=dev)
        for i, m in enumerate(mods):
            out[..., i, :] = torch.randint(0, m, shape_prefix + (n,), generator=g, device=dev, dtype=torch.int64)
        return out

    a = rand_rows((B, 2), MODULI[:k])
    b = rand_rows((B, 2), MODULI[:k])
    rlk = rand_rows((k, 2), MODULI)
    out = torch.zeros((B, 2, k, n), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.multiply_relin(a, b, rlk, out, B, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # keep the GPU under the same load until the sampler is producing lines, then time
    t_wait = time.time()
    while rank == 0 and not sampler.samples and time.time() - t_wait < 3.0:
        step()
        torch.cuda.synchronize()
    sampler.mark = len(sampler.samples)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    if dist is not None:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1000.0)

    # ---- e2e: host buffers through the C ABI (H2D + compute + D2H in the timed region) ----
    ct_bytes = 2 * k * n * 8
    ah = torch.empty((B, 2, k, n), dtype=torch.int64).pin_memory()
    bh = torch.empty((B, 2, k, n), dtype=torch.int64).pin_memory()
    oh = torch.empty((B, 2, k, n), dtype=torch.int64).pin_memory()
    ah.copy_(a)
    bh.copy_(b)
    e2e_steps = max(2, min(args.steps, 5))
    ctx.multiply_relin_host(ah, bh, rlk, oh, B)  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.multiply_relin_host(ah, bh, rlk, oh, B)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        ts = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = world * B * e2e_steps / e2e_s
    same = bool(torch.equal(oh.to(dev), out))
    pack_num = 6 if (max(MODULI[:k]) < 2 ** 48 and os.environ.get("B200_HOST_PACK", "0") not in ("", "0")) else 8

    # ---- roofline of the dominant kernel: batched forward NTT, 4096 polys x 4 residues (1 GiB slab > L2) ----
    roof = None
    cpu = None
    if rank == 0:
        peak, peak_src = peaks()
        slab = rand_rows((NTT_POLYS,), MODULI[:k])
        for _ in range(3):
            ctx.ntt_forward(slab, NTT_POLYS, stream=stream)
            ctx.ntt_inverse(slab, NTT_POLYS, stream=stream)
        torch.cuda.synchronize()
        reps = 10
        ef0, ef1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fwd_ms = 0.0
        inv_ms = 0.0
        for _ in range(reps):
            ef0.record()
            ctx.ntt_forward(slab, NTT_POLYS, stream=stream)
            ef1.record()
            torch.cuda.synchronize()
            fwd_ms += ef0.elapsed_time(ef1)
            ef0.record()
            ctx.ntt_inverse(slab, NTT_POLYS, stream=stream)
            ef1.record()
            torch.cuda.synchronize()
            inv_ms += ef0.elapsed_time(ef1)
        fwd_ms /= reps
        inv_ms /= reps
        alg_bytes = 16 * n * NTT_POLYS * k  # 16*n bytes per residue NTT
        achieved = alg_bytes / (fwd_ms / 1000.0) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "ntt_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "ntt_fp_kernel<13, fwd, 256> (4096 polys x 4 residues, n=8192; FP64 butterflies)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                "inverse_achieved": alg_bytes / (inv_ms / 1000.0) / 1e9,
                "mul_relin_algorithmic_gbs": (8 * n * k * 6) * value / world / 1e9}
        del slab
        if not args.no_cpu and world == 1:  # the CPU baseline is an N=1 leg (the reference arm covers every N)
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))
                cores = os.cpu_count() or 1
                iters = 32
                rate, secs = cpu_reference_rate(cores, iters)
                rate1, secs1 = cpu_reference_rate(1, 16)
                cpu = {"value": rate, "unit": "ops/s", "cores": cores, "kind": "reference",
                       "sample": f"{cores} threads x {iters} multiply+relinearize_inplace (uniform-random words), {secs:.2f}s",
                       "single_thread_ops_per_s": rate1}
            except Exception as ex:  # reference .so missing on this box
                cpu = {"value": None, "unit": "ops/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": "ciphertext mul+relin/sec", "value": value, "unit": "ops/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "Evaluator::multiply + relinearize, n=8192, BFVDefault 218-bit chain (k=4 + special), "
                                   f"batch={B} ct pairs per GPU", "batch_per_gpu": B,
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                       "l2": "inputs (1 GiB per GPU) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "ops/s",
                    # bytes that actually cross PCIe per step (with B200_HOST_PACK=1 the library narrows each residue word
                    # to 6 bytes on the host and widens it again on the device; off by default: measured slower)
                    "h2d_bytes_per_step": 2 * B * ct_bytes * pack_num // 8, "d2h_bytes_per_step": B * ct_bytes * pack_num // 8,
                    "host_buffer_bytes_in_per_step": 2 * B * ct_bytes, "host_buffer_bytes_out_per_step": B * ct_bytes,
                    "transfer": "6-byte packed residues" if pack_num == 6 else "8-byte words",
                    "steps": e2e_steps, "matches_device_path": same, "host_numa_node": numa_node},
            "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
