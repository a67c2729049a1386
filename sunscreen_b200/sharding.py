"""Batch sharding across GPUs: independent ciphertext operations are split by contiguous slices of
the batch index — item b goes to rank floor(b * world / batch) — keys and tables are replicated per rank, and a
collective is only used to scatter inputs from / gather outputs to rank 0.  No collective inside the compute.

One process per GPU (torchrun); backend "nccl" on GPUs, "gloo" in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous slice [lo, hi) of the batch owned by `rank` (sizes differ by at most one)."""
    lo = (batch * rank) // world
    hi = (batch * (rank + 1)) // world
    return lo, hi


def scatter_batch(full, batch, item_shape, dtype, device, src=0):
    """Rank `src` holds `full` [batch, *item_shape]; every rank returns its own slice (padded scatter)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_bounds(batch, rank, world)
    width = max(shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world))
    recv = torch.empty((width,) + tuple(item_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = shard_bounds(batch, r, world)
            c = torch.zeros((width,) + tuple(item_shape), dtype=dtype, device=device)
            c[: b - a] = full[a:b].to(device)
            chunks.append(c)
    dist.scatter(recv, chunks, src=src)
    return recv[: hi - lo].contiguous()


def gather_batch(local, batch, item_shape, dtype, device, dst=0):
    """Inverse of scatter_batch: rank `dst` gets [batch, *item_shape], others None."""
    rank, world = dist.get_rank(), dist.get_world_size()
    width = max(shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world))
    send = torch.zeros((width,) + tuple(item_shape), dtype=dtype, device=device)
    send[: local.shape[0]] = local
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((batch,) + tuple(item_shape), dtype=dtype, device=device)
    for r in range(world):
        a, b = shard_bounds(batch, r, world)
        out[a:b] = bufs[r][: b - a]
    return out


def sharded_multiply_relin(ctx, a_full, b_full, rlk, batch, device, level=None, to_ptr=None):
    """multiply+relinearize of `batch` pairs held on rank 0, computed on all ranks.  `ctx` is this rank's B200Context,
    `rlk` this rank's replica of the relinearization key.  Returns the gathered result on rank 0."""
    k, n = ctx.k(level), ctx.n
    shape = (2, k, n)
    a = scatter_batch(a_full, batch, shape, torch.int64, device)
    b = scatter_batch(b_full, batch, shape, torch.int64, device)
    out = torch.zeros_like(a)
    if a.shape[0]:
        ctx.multiply_relin(a, b, rlk, out, a.shape[0], level=level)
    if device != "cpu" and torch.cuda.is_available():
        torch.cuda.synchronize()
    return gather_batch(out, batch, shape, torch.int64, device)
