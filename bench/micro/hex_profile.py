"""Where does the time of the N = 100 hexagonal solve go?  cProfile of the resident solve + the library's own per-call time."""
import cProfile
import os
import pstats
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
for _ in range(2):
    rb.slab.reset()
    solver.solveGraph(solver.generateGraph_Hexagonal(N=N), rb, gibbs_iters=3, seed=1)
best = 1e9
for _ in range(5):
    rb.slab.reset()
    fg = solver.generateGraph_Hexagonal(N=N)
    t0 = time.perf_counter(); solver.solveGraph(fg, rb, gibbs_iters=3, seed=1); best = min(best, time.perf_counter() - t0)
print(f"resident solve N={N}: {best*1e3:.2f} ms, kernel launches per solve: ", end="")
l0 = ctx.kernel_launches; rb.slab.reset(); solver.solveGraph(solver.generateGraph_Hexagonal(N=N), rb, gibbs_iters=3, seed=1); print(ctx.kernel_launches - l0)
pr = cProfile.Profile()
rb.slab.reset()
fg = solver.generateGraph_Hexagonal(N=N)
pr.enable(); solver.solveGraph(fg, rb, gibbs_iters=3, seed=1); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
