"""One warm-up and one measured resident hex solve (N from argv): run under `ncu --metrics gpu__time_duration.sum --csv` to list the
device time of every kernel of the second solve (the per-call floor of the small-batch path)."""
import os
import sys
import time
from importlib import import_module

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

cb = load_package()
solver = import_module("caesar_jl_b200.solver")
ctx = cb.Context(0, "fp32")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rb = solver.DeviceResidentBackend(cb, ctx)
for k in range(2):
    rb.slab.reset()
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    solver.solveGraph(solver.generateGraph_Hexagonal(N=N), rb, gibbs_iters=3, seed=1)
    print(f"solve {k}: {(time.perf_counter() - t0) * 1e3:.2f} ms, {ctx.kernel_launches - l0} launches")
