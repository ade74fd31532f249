import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import oracle
cb = load_package()
ctx = cb.Context(0, "fp32")
rng = np.random.default_rng(0)
for (D, N, circ) in [(2, 64, (0,)*6), (2, 128, (0,)*6), (2, 100, (0,)*6), (2, 300, (0,)*6), (4, 300, (0, 0, 0, 1, 1, 1)), (3, 1024, (0,)*6), (2, 4096, (0,)*6)]:
    dens = []
    for j in range(D):
        pts = rng.normal(size=(N, 6)) * 0.5 + rng.normal(size=6) * 0.2
        dens.append((pts, np.full(6, 0.3), None))
    pts, lab = cb.product_points(dens, circ, 512, seed=3, ctx=ctx)
    opts, olab = oracle.product(dens, 512, seed=3, circular=list(circ), ubits=24)
    same = (lab == olab).all(axis=1)
    print(D, N, circ, "label agreement", same.mean(), "finite", np.isfinite(pts).all(), "per-message", [(lab[:, j] == olab[:, j]).mean().round(3) for j in range(D)])
