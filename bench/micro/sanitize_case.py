"""Smallest run that touches every kernel family once (for compute-sanitizer --tool memcheck): ragged product shapes (d = 3 Pose2,
d = 5, d = 6 with the tensor-core leaf pass and a partial last tile), bandwidth searches (single CTA, cluster, multi-launch), mmd,
SE batch, evaluate, sample, closed-form convolutions, residual."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

cb = load_package()
rng = np.random.default_rng(0)
for prec in ("fp32", "fp64"):
    ctx = cb.Context(0, prec)
    for d, circ, N in ((3, (0, 0, 1), 130), (5, (0, 0, 0, 1, 1), 300), (6, (0, 0, 0, 1, 1, 1), 257)):
        dens = [(rng.normal(size=(N + 3 * j, d)) * 0.5, np.full(d, 0.3)) for j in range(3)]
        pts, lab = cb.product_points(dens, circ, 150, seed=1, ctx=ctx)
        assert np.isfinite(pts).all()
    for N in (100, 300, 1100):
        bw = cb.kde_bandwidth(rng.normal(size=(N, 3)), "Pose2", ctx=ctx)
        assert (bw > 0).all()
    a, b = rng.normal(size=(700, 3)), rng.normal(size=(300, 3))
    cb.mmd(a, b, 0.5, ctx=ctx)
    cb.mmd_sums_shard(a, b, 0.5, 1, 3, ctx=ctx)
    cb.mmd_se_batch(a, b, 0.5, rng.normal(size=(3, 6)) * 0.1, ctx=ctx)
    m = cb.ManifoldKernelDensity("Point3", a, [0.2, 0.2, 0.2], weights=rng.uniform(0.1, 1, size=700))
    cb.evaluate(m, b, ctx=ctx)
    cb.sample(m, 50, seed=2, ctx=ctx)
    src = np.c_[rng.normal(size=(40, 3)), rng.normal(size=(40, 3)) * 0.3]
    cb.approxConvBatch([dict(kind=4, reverse=False, N=64, mean=np.zeros(6), chol=0.1 * np.eye(6), src=src)], seed=3, ctx=ctx)
    cb.se_residual_batch(3, src, src, src, ctx=ctx)
print("sanitize case ok")
