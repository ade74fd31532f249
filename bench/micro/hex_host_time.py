"""Host time per call type of the resident hex solve (no device waits inside: everything is stream-ordered)."""
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
acc = {}


class Timed(solver.DeviceResidentBackend):
    def _t(self, name, f, *a):
        t0 = time.perf_counter()
        r = f(*a)
        e = acc.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += time.perf_counter() - t0
        return r

    def manikde_batch(self, *a):
        return self._t("manikde", super().manikde_batch, *a)

    def products(self, *a):
        return self._t("products", super().products, *a)

    def approx_conv_batch(self, *a):
        return self._t("conv", super().approx_conv_batch, *a)

    def finish(self, *a):
        return self._t("finish(waits)", super().finish, *a)


rb = Timed(cb, ctx)
for k in range(4):
    acc.clear()
    rb.slab.reset()
    fg = solver.generateGraph_Hexagonal(N=N)
    t0 = time.perf_counter()
    solver.solveGraph(fg, rb, gibbs_iters=3, seed=1)
    tot = time.perf_counter() - t0
print(f"N={N} solve {tot * 1e3:.2f} ms")
for k, (n, t) in acc.items():
    print(f"  {k:14s} calls {n:3d}  total {t * 1e3:6.3f} ms  per call {t / n * 1e6:6.1f} us")
print(f"  harness (python outside the calls): {(tot - sum(t for _, t in acc.values())) * 1e3:.3f} ms")
