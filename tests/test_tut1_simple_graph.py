"""The reference's ICRA tutorial 1 test (R:test/ICRA2022_tests/simple_graph.jl) on the caller harness: a Pose2 square driven by four
odometry factors, one landmark sighted from the first and (after the second solve) the last pose, with the reference's own
assertions: point estimates within 0.2 (landmark 0.1), belief means within 0.2 in the norm of SE(2) as an ArrayPartition
(translation difference and Frobenius difference of the rotation matrices)."""
from importlib import import_module

import numpy as np
import pytest

from __graft_entry__ import build, load_package

ODO = (np.array([1.0, 0.0, np.pi / 2]), np.diag([0.1, 0.1, 0.01]) ** 2)   # :22


def _se2_isapprox(ppe, xy, theta, atol):
    """isapprox(ArrayPartition(xy, RotMatrix2(theta)), mean(X); atol)  (:58-60): norm of the partition difference"""
    dth = ppe[2] - theta
    return np.sqrt((ppe[0] - xy[0]) ** 2 + (ppe[1] - xy[1]) ** 2 + 8.0 * np.sin(0.5 * dth) ** 2) <= atol


def _vec_isapprox(ppe, want, atol):
    e = np.asarray(ppe, dtype=float) - np.asarray(want, dtype=float)
    if len(e) == 3:
        e[2] = (e[2] + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(e) <= atol


def _run_reference_test(solver, backend, seed):
    ppe = lambda v: solver.getPPE(fg.variables[v].belief)
    fg = solver.FactorGraph(N=100)
    fg.addVariable("x0", "Pose2")
    fg.addFactor(["x0"], solver.PriorPose2(np.zeros(3), np.diag([0.05, 0.05, 0.01]) ** 2))   # :14-15
    fg.addVariable("x1", "Pose2")
    fg.addFactor(["x0", "x1"], solver.Pose2Pose2(*ODO))                                      # :21-22
    solver.solveGraph(fg, backend, seed=seed)                                                  # :25
    assert _vec_isapprox(ppe("x0"), [0, 0, 0], 0.2) and _vec_isapprox(ppe("x1"), [1, 0, np.pi / 2], 0.2)   # :27-28
    fg.addVariable("l1", "Point2")
    fg.addFactor(["x0", "l1"], solver.Pose2Point2BearingRange((0.0, 0.03), (0.5, 0.1)))      # :36-37
    for i in (2, 3, 4):                                                                        # :39-46
        fg.addVariable(f"x{i}", "Pose2")
        fg.addFactor([f"x{i - 1}", f"x{i}"], solver.Pose2Pose2(*ODO))

    def checks():
        assert _vec_isapprox(ppe("x0"), [0, 0, 0], 0.2) and _vec_isapprox(ppe("x1"), [1, 0, np.pi / 2], 0.2)   # :53-54
        assert _vec_isapprox(ppe("l1"), [0.5, 0], 0.1)                                                      # :55
        assert _se2_isapprox(ppe("x2"), [1, 1], np.pi, 0.2)                                                 # :57-59
        assert _se2_isapprox(ppe("x3"), [0, 1], -np.pi / 2, 0.2)                                            # :61-63
        assert _se2_isapprox(ppe("x4"), [0, 0], 0.0, 0.2)                                                   # :65-67

    solver.solveGraph(fg, backend, seed=seed + 1)                                              # :50
    checks()
    fg.addFactor(["x4", "l1"], solver.Pose2Point2BearingRange((0.0, 0.03), (0.5, 0.1)))      # :75-76: loop closure
    solver.solveGraph(fg, backend, seed=seed + 2)                                              # :78
    checks()                                                                                   # :80-98


@pytest.mark.parametrize("seed", [0, 5, 9])
def test_tutorial1_on_oracle_backend(seed):
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    from oracle.solver_backend import OracleBackend
    _run_reference_test(solver, OracleBackend(), seed)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_tutorial1_on_device(precision):
    build()
    cb = load_package()
    solver = import_module("caesar_jl_b200.solver")
    ctx = cb.Context(0, precision)
    for backend in (solver.DeviceBackend(cb, ctx), solver.DeviceResidentBackend(cb, ctx)):
        for seed in (0, 5, 9):
            _run_reference_test(solver, backend, seed)
