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


def mmd(a, b, bw, ctx, dist=None, shard_fn=None):
    """`mmd(M, a, b, N, M; bw)` from HOST arrays on every rank of a multi-GPU job: both clouds are uploaded to each GPU (they are
    KBs to MBs), every rank evaluates its share of the symmetric tile list (cb2_mmd_sums_shard), and the three FP64 partial sums
    are all-reduced over NCCL.  Every rank returns the same value."""
    import torch
    from . import mmd_sums_shard, mmd_combine
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    # shard_fn(a, b, bw, shard, nshards) -> the three partial sums of that shard (tests: a CPU stand-in for the device call)
    part = np.asarray(shard_fn(a, b, bw, rank, world), dtype=np.float64) if shard_fn else mmd_sums_shard(a, b, bw, rank, world, ctx=ctx)
    if dist is not None and world > 1:
        t = torch.from_numpy(part)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        part = t.cpu().numpy()
    return mmd_combine(part, len(a), len(b))


def product_sharded_dev(descs_fn, Nout, d, D, dist, torch, ctx_call):
    """One large product with its OUTPUT SAMPLES split over the ranks, everything on the device: rank r runs the chains
    [s0, s1) of the product into its slice of the full output tensors (Philox sample index = global index, so the result does
    not depend on the world size) and the slices are all-gathered over NCCL.  descs_fn(s0, n, pts_ptr, lab_ptr) builds the
    device-pointer descriptor of that slice; ctx_call(desc) launches it.  Returns (pts Nout x d, labels Nout x D) CUDA tensors,
    identical on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    cap = -(-Nout // world)
    pts = torch.zeros((world * cap, d), dtype=torch.float64, device="cuda")
    lab = torch.zeros((world * cap, D), dtype=torch.int32, device="cuda")
    s0, s1 = min(rank * cap, Nout), min((rank + 1) * cap, Nout)
    if s1 > s0:
        ctx_call(descs_fn(s0, s1 - s0, pts[rank * cap:].data_ptr(), lab[rank * cap:].data_ptr()))
    if world > 1:
        dist.all_gather_into_tensor(pts, pts[rank * cap:(rank + 1) * cap].clone())
        dist.all_gather_into_tensor(lab, lab[rank * cap:(rank + 1) * cap].clone())
    return pts[:Nout], lab[:Nout]


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
