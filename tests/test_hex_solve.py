"""BASELINE config C2: hexagonal Pose2 SLAM (generateGraph_Hexagonal, N=100 particles) solved through the batched
device product path, compared with the same caller harness running on the CPU oracle."""
import numpy as np
import pytest

from __graft_entry__ import build, load_package

HEX_TRUTH = {"x0": (0, 0, 0), "x1": (10, 0, np.pi / 3), "x2": (15, 8.660254, 2 * np.pi / 3), "x3": (10, 17.320508, np.pi),
             "x4": (0, 17.320508, -2 * np.pi / 3), "x5": (-5, 8.660254, -np.pi / 3), "x6": (0, 0, 0), "l1": (20, 0)}


def _errors(solver, fg):
    err = {}
    for k, t in HEX_TRUTH.items():
        ppe = solver.getPPE(fg.variables[k].belief)
        e = ppe - np.array(t, dtype=float)
        if len(t) == 3:
            e[2] = (e[2] + np.pi) % (2 * np.pi) - np.pi
        err[k] = e
    return err


def test_hex_solve_on_oracle_backend_recovers_hexagon():
    """The caller harness itself (CPU oracle products): poses within the reference's own loose bands
    (R:test/ICRA2022_tests/simple_graph.jl style PPE checks)."""
    from importlib import import_module
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    from oracle.solver_backend import OracleBackend
    fg = solver.generateGraph_Hexagonal(N=100)
    st = solver.solveGraph(fg, OracleBackend(), gibbs_iters=3, seed=1)
    err = _errors(solver, fg)
    for k, e in err.items():
        assert np.abs(e[:2]).max() < 1.5, (k, e)
        if len(e) == 3:
            assert abs(e[2]) < 0.2, (k, e)
    assert st["products"] >= 3 * 7 and set(st["shapes"]) <= {"D=2,N=100,d=3", "D=3,N=100,d=3", "D=2,N=100,d=2"}


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_hex_solve_on_device_matches_oracle_solve(precision):
    from importlib import import_module
    build()
    cb = load_package()
    solver = import_module("caesar_jl_b200.solver")
    from oracle.solver_backend import OracleBackend
    ctx = cb.Context(0, precision)
    fg = solver.generateGraph_Hexagonal(N=100)
    st = solver.solveGraph(fg, solver.DeviceBackend(cb, ctx), gibbs_iters=3, seed=1)
    err = _errors(solver, fg)
    for k, e in err.items():
        assert np.abs(e[:2]).max() < 1.5, (k, e)
        if len(e) == 3:
            assert abs(e[2]) < 0.2, (k, e)
    # same harness, same host RNG, products on the oracle: beliefs agree under the reference's mmd tolerance
    fgo = solver.generateGraph_Hexagonal(N=100)
    solver.solveGraph(fgo, OracleBackend(ubits=53 if precision == "fp64" else 24), gibbs_iters=3, seed=1)
    for k in HEX_TRUTH:
        a, b = fg.variables[k].belief, fgo.variables[k].belief
        assert cb.mmd(a.pts[:, :2], b.pts[:, :2], bw=0.001, ctx=ctx) < 0.01, k  # insufficient_data.jl's tightest atol
    assert st["products"] == 24 and st["product_calls"] <= 3 * 6
    if precision == "fp64":
        # same Philox draws for the factor convolutions and the Gibbs chains, FP64 arithmetic on both sides: the whole
        # three-sweep solve (24 products, 57 bandwidth searches, 57 convolutions) reproduces the oracle's beliefs
        for k in HEX_TRUTH:
            a, b = fg.variables[k].belief, fgo.variables[k].belief
            assert np.isclose(a.pts, b.pts, atol=1e-6).all(axis=1).mean() >= 0.97, k
            assert np.allclose(a.bw, b.bw, rtol=1e-3), k


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("N", [100, 1000, 1500])
def test_hex_solve_with_device_resident_beliefs_equals_the_host_pointer_solve(precision, N):
    """proposals, bandwidths and products as `*_dev` calls on beliefs that never leave the device must give the
    beliefs of the host-pointer path bit for bit (same kernels, same Philox counters).  N = 1500 takes the multi-launch
    bandwidth search (N > 1024) and N not a multiple of the chain block."""
    from importlib import import_module
    build()
    cb = load_package()
    solver = import_module("caesar_jl_b200.solver")
    ctx = cb.Context(0, precision)
    fa = solver.generateGraph_Hexagonal(N=N)
    sa = solver.solveGraph(fa, solver.DeviceBackend(cb, ctx), gibbs_iters=2, seed=5)
    fb = solver.generateGraph_Hexagonal(N=N)
    L = cb._capi.lib()
    launches0 = L.cb2_kernel_launches(ctx._h)
    sb = solver.solveGraph(fb, solver.DeviceResidentBackend(cb, ctx), gibbs_iters=2, seed=5)
    assert L.cb2_kernel_launches(ctx._h) > launches0
    assert sa["products"] == sb["products"] and sa["shapes"] == sb["shapes"]
    for k in HEX_TRUTH:
        a, b = fa.variables[k].belief, fb.variables[k].belief
        assert np.array_equal(a.pts, b.pts), k
        assert np.array_equal(a.bw, b.bw), k


def test_solve_tree_verb_on_oracle_backend():
    """ZMQ verb `solveTree` (R:ext/services/session.jl:133-140): reply fields and a solved graph."""
    from importlib import import_module
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    ser = import_module("caesar_jl_b200.serialization")
    from oracle.solver_backend import OracleBackend
    fg = solver.generateGraph_Hexagonal(N=60)
    resp = ser.solveTree({}, fg, {"request": "solveTree"}, backend=OracleBackend(), gibbs_iters=2, seed=3)
    assert resp["status"] == "OK" and resp["durationSec"] > 0 and set(resp) == {"startTime", "endTime", "durationSec", "status"}
    assert np.abs(solver.getPPE(fg.variables["x3"].belief)[:2] - [10, 17.32]).max() < 2.0
