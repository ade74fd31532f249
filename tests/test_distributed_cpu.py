"""World-size-2 gloo tests (CPU) of the multi-GPU sharding logic: row-sharded mmd with an all-reduce of three
partial sums, output-sample-sharded product with an all-gather, round-robin variable batches.  The per-rank compute
is played by the oracle here; on the GPU box the same helpers wrap the C-ABI calls (tests/test_gpu_parity.py checks
that the device partial sums / sample_offset slices equal the unsharded result)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import oracle
from __graft_entry__ import load_package


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cbd = load_package().distributed if hasattr(load_package(), "distributed") else None
    from importlib import import_module
    cbd = import_module("caesar_jl_b200.distributed")
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(301, 3)), rng.normal(size=(250, 3)) + 0.4
    val, sums = cbd.mmd_sharded(a, b, 0.6, lambda a, b, bw, ra, rb: oracle.mmd_rows(a, b, bw, ra, rb), dist)
    dens = [(rng.normal(size=(64, 2)), np.array([0.3, 0.3])), (rng.normal(size=(50, 2)) + 0.5, np.array([0.2, 0.4]))]
    pf = lambda dens, circ, n, seed, niter, stream, sample_offset: oracle.product(dens, n, seed, circular=list(circ), niter=niter,
                                                                                  stream=stream, sample_offset=sample_offset)
    pts, lab = cbd.product_sharded(dens, (0, 0), 101, 5, pf, dist, stream=3)
    mine = cbd.assign_round_robin(7, rank, world)
    # the host-array entry point of the multi-GPU mmd (tile shards on the device; here: interleaved rows on the oracle)
    def shard(a, b, bw, r, w):
        s = np.zeros(3)
        for i in range(r, len(a), w):
            s[:2] += oracle.mmd_rows(a, b, bw, (i, i + 1), (0, 0))[:2]
        for i in range(r, len(b), w):
            s[2] += oracle.mmd_rows(a, b, bw, (0, 0), (i, i + 1))[2]
        return s
    val2 = cbd.mmd(a, b, 0.6, None, dist, shard_fn=shard)
    q.put((rank, val, sums, pts, lab, mine, val2))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_sharding_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(301, 3)), rng.normal(size=(250, 3)) + 0.4
    want, wsums = oracle.mmd(a, b, 0.6, return_sums=True)
    dens = [(rng.normal(size=(64, 2)), np.array([0.3, 0.3])), (rng.normal(size=(50, 2)) + 0.5, np.array([0.2, 0.4]))]
    wpts, wlab = oracle.product(dens, 101, 5, circular=[0, 0], stream=3)
    for rank, val, sums, pts, lab, mine, val2 in res:
        assert np.isclose(val, want, rtol=1e-12) and np.allclose(sums, wsums, rtol=1e-12)
        assert np.isclose(val2, want, rtol=1e-10, atol=1e-14)
        assert np.array_equal(pts, wpts) and np.array_equal(lab, wlab)  # identical for any world size
    assert sorted(res[0][5] + res[1][5]) == list(range(7))


def test_shard_range_covers_everything():
    from importlib import import_module
    load_package()
    cbd = import_module("caesar_jl_b200.distributed")
    for n in (0, 1, 7, 100000):
        for w in (1, 2, 3, 8):
            r = [cbd.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
