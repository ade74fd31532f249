"""One warm-up and measured resident hex solves (N from argv).  Prints the solve time, the launch count and how long the host
needed to enqueue everything (the time until the final download starts waiting).  Run under
`ncu --metrics gpu__time_duration.sum --csv` to list the device time of every kernel (bench/micro/timeline_summary.py)."""
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
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2


class Timed(solver.DeviceResidentBackend):
    def finish(self, fg):
        self.t_enqueued = time.perf_counter()
        super().finish(fg)


rb = Timed(cb, ctx)
best = (1e9, 0, 0)
for k in range(reps):
    rb.slab.reset()
    l0 = ctx.kernel_launches
    fg = solver.generateGraph_Hexagonal(N=N)
    t0 = time.perf_counter()
    solver.solveGraph(fg, rb, gibbs_iters=3, seed=1)
    t1 = time.perf_counter()
    if k and t1 - t0 < best[0]:
        best = (t1 - t0, rb.t_enqueued - t0, ctx.kernel_launches - l0)
print(f"N={N}: best solve {best[0] * 1e3:.2f} ms, host enqueue {best[1] * 1e3:.2f} ms, {best[2]} launches")
