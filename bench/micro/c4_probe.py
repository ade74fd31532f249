"""Stage timers of the C4-shaped batch (256 Pose2 products, D = 3, N = N_out = 1000) through the host-pointer call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

cb = load_package()
ctx = cb.Context(0, "fp32")
L = cb._capi.lib()
corners = np.array([[0, 0, 0.0], [10, 0, 1.5], [10, 10, 3.0], [0, 10, -1.5]])


def group(v):
    r = np.random.default_rng(4000 + v)
    g = []
    for j in range(3):
        cs = corners.copy()
        cs[1:] += r.normal(size=(3, 3)) * [3.0, 3.0, 0.5] * (j + 1)
        pts = cs[r.integers(0, 4, size=1000)] + r.normal(size=(1000, 3)) * [0.4, 0.4, 0.1]
        pts[:, 2] = (pts[:, 2] + np.pi) % (2 * np.pi) - np.pi
        g.append(cb.ManifoldKernelDensity("Pose2", pts, 0.53 * pts.std(0) * 1000 ** -0.2))
    return g


groups = [group(v) for v in range(256)]
for it in range(3):
    t0 = time.perf_counter()
    cb.manifoldProducts(groups, "Pose2", N=1000, seed=2 + it, ctx=ctx)
    dt = time.perf_counter() - t0
    st = {k: round(L.cb2_last_stage_ms(ctx._h, k.encode()), 2) for k in ("h2d", "tree", "gibbs", "bw", "d2h")}
    print(f"call {it}: {dt * 1e3:.1f} ms host;", st, "sum", round(sum(st.values()), 1))
