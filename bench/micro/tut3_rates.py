"""How often the reference's 47 tutorial-3 checks (R:test/ICRA2022_tests/insufficient_data.jl:143-248) hold for the harness on the device:
share of checks over the seeds of tests/test_tut3_range_slam.py, whole runs that pass, the checks that miss most."""
import os
import sys
from importlib import import_module

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import test_tut3_range_slam as T

cb = load_package()
solver = import_module("caesar_jl_b200.solver")
z = np.load(os.path.join(ROOT, "tests", "golden", "tut3_beliefs.npz"))
tut3 = {str(n): (z["pts"][i], z["bw"][i]) for i, n in enumerate(z["names"])}
seeds = range(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
for precision in ("fp32", "fp64"):
    ctx = cb.Context(0, precision)
    dists = [T._solve_five_stages(solver, solver.DeviceResidentBackend(cb, ctx), s, tut3) for s in seeds]
    checks = [((st, v), a) for st, tol in T.ATOL_ALL.items() for v, a in tol.items()]
    ok = np.array([[d[k] < a for k, a in checks] for d in dists])
    miss = sorted(((int((~ok[:, i]).sum()), checks[i][0]) for i in range(len(checks)) if not ok[:, i].all()), reverse=True)
    print(f"{precision}: {ok.sum()} of {ok.size} checks hold over {len(dists)} solves, {int(ok.all(axis=1).sum())} solves pass all 47; misses: {miss[:6]}")
