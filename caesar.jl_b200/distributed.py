"""Multi-GPU sharding of the hot path.  One process per GPU; torch.distributed is the plumbing
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  Only the steps that really exchange data use a collective:

  mmd / evaluate : both clouds replicated, ROWS of the three pair sums sharded, 3 FP64 partials all-reduced(sum)
  one product    : messages replicated, OUTPUT SAMPLES split, Philox sample index = global index so the result is
                   identical for any world size; points (+labels) all-gathered
  many products  : independent variables / ready cliques dealt round-robin to ranks, no data-path collective
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous balanced split of range(n): returns (begin, end) of `rank`."""
    base, rem = divmod(n, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def assign_round_robin(n_items, rank, world):
    return list(range(rank, n_items, world))


def mmd_sharded(a, b, bw, sums_fn, dist=None, device=None):
    """Row-sharded mmd.  sums_fn(a, b, bw, rows_a, rows_b) -> np.array([aa, ab, bb]) restricted to those rows
    (caesar_jl_b200.mmd_sums on the GPU; the tests pass the oracle).  All-reduces the three partial sums."""
    import torch
    N, M = len(a), len(b)
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    ra, rb = shard_range(N, rank, world), shard_range(M, rank, world)
    part = np.asarray(sums_fn(a, b, bw, ra, rb), dtype=np.float64)
    if dist is not None and world > 1:
        t = torch.from_numpy(part.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        part = t.cpu().numpy()
    return part[0] / (N * N) - 2.0 * part[1] / (N * M) + part[2] / (M * M), part


def product_sharded(dens, circular, Nout, seed, product_fn, dist=None, device=None, stream=0, niter=1):
    """One large product with its output samples split across ranks and all-gathered.
    product_fn(dens, circular, n, seed=, niter=, stream=, sample_offset=) -> (pts n x d, labels n x D)."""
    import torch
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    s0, s1 = shard_range(Nout, rank, world)
    pts, lab = product_fn(dens, circular, s1 - s0, seed=seed, niter=niter, stream=stream, sample_offset=s0)
    if dist is None or world == 1:
        return pts, lab
    d, D = pts.shape[1], lab.shape[1]
    cap = -(-Nout // world)  # all_gather needs equal shapes: pad to the largest shard
    buf = np.zeros((cap, d + D))
    buf[: s1 - s0, :d] = pts
    buf[: s1 - s0, d:] = lab
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    parts = []
    for r in range(world):
        b, e = shard_range(Nout, r, world)
        parts.append(out[r].cpu().numpy()[: e - b])
    full = np.concatenate(parts, axis=0)
    return full[:, :d].copy(), full[:, d:].astype(np.int32)
