"""The reference's own fixture test, as far as this repository's scope reaches: R:test/ICRA2022_tests/insufficient_data.jl solves the
range-only SLAM graph of ICRA tutorial 3 stage by stage and requires `mmd(stored belief, solved belief) < atol` against the 44
beliefs UPSTREAM stored under test_data/tut3 (tests/golden/tut3_beliefs.npz holds them byte for byte).  Here the same graph is solved
by the caller harness (solver.py: proposals -> manikde! -> manifoldProduct -> manikde!) and held to the reference's tolerances with
the reference's mmd (bw = 0.001): its first two checkpoints (:143-171) with the harness's default three sweeps per solve, and ALL
FIVE checkpoints (47 checks, :143-248) with one sweep per solve -- upstream's stored beliefs of the later stages are partially
converged (x3_3 still spreads 86 m), as one upward / downward pass of IIF's Bayes tree leaves them, and three sweeps of the whole
graph converge further than that.  This pins product + bandwidth + mmd, end to end, to upstream-produced numbers."""
from importlib import import_module

import numpy as np
import pytest

import oracle
from __graft_entry__ import build, load_package

# insufficient_data.jl:143-144 (stage 1) and :160-166 (stage 2)
ATOL = {1: {"l1": 0.01, "l2": 0.01, "l3": 0.1, "x0": 0.1},
        2: {"l1": 0.01, "l2": 0.01, "l3": 0.3, "l4": 0.3, "x0": 0.2, "x1": 0.4, "x2": 0.4}}


def _solve_two_stages(solver, backend, seed):
    fg = solver.generateGraph_Tut3(N=100)
    out = {}
    for stage in (1, 2):
        if stage == 2:  # :151-152
            solver.tut3_vehicle_drives(fg, "x0", "x1")
            solver.tut3_vehicle_drives(fg, "x1", "x2")
        solver.solveGraph(fg, backend, gibbs_iters=3, seed=seed + stage)
        out[stage] = {v: np.array(_pts(fg.variables[v].belief)) for v in ATOL[stage]}
    return fg, out


# all five checkpoints of the reference test (:143-248)
ATOL_ALL = {1: ATOL[1], 2: ATOL[2],
            3: {"l1": 0.01, "l2": 0.01, "l3": 0.5, "l4": 0.5, "x0": 0.2, "x1": 0.3, "x2": 0.3, "x3": 0.5, "x4": 0.5},
            4: {"l1": 0.01, "l2": 0.01, "l3": 0.6, "l4": 0.6, "x0": 0.4, "x1": 0.7, "x2": 0.7, "x3": 0.7, "x4": 0.7, "x5": 0.7, "x6": 0.7},
            5: {"l1": 0.01, "l2": 0.01, "l3": 0.8, "l4": 0.8, "x0": 0.7, "x1": 0.7, "x2": 0.7, "x3": 0.7, "x4": 0.7, "x5": 0.8, "x6": 0.8,
                "x7": 0.7, "x8": 0.7}}
SEEDS_ALL = tuple(range(8))


def _solve_five_stages(solver, backend, seed, tut3):
    """The whole reference test with ONE sweep per `solveGraph!`: IIF's solve is one upward and one downward pass over the Bayes tree
    (its `gibbsIters = 3` act inside a clique), which a single sweep over the graph resembles more than three sweeps of the whole
    graph as one clique do.  Returns {(stage, variable): mmd to the stored belief}."""
    fg = solver.generateGraph_Tut3(N=100)
    dist = {}
    for stage in range(1, 6):
        if stage > 1:  # :151-152, :174-175, :198-199, :224-225
            a = 2 * (stage - 1) - 2
            solver.tut3_vehicle_drives(fg, f"x{a}", f"x{a + 1}")
            solver.tut3_vehicle_drives(fg, f"x{a + 1}", f"x{a + 2}")
        solver.solveGraph(fg, backend, gibbs_iters=1, seed=seed + stage)
        for v in ATOL_ALL[stage]:
            dist[(stage, v)] = oracle.mmd(tut3[f"{v}_{stage}"][0], _pts(fg.variables[v].belief), 0.001)
    return dist


def _check_all(dists):
    """The reference test is one stochastic realisation per run and its later tolerances sit close to what a run produces (on the
    oracle 547 of 564 checks hold over twelve seeds, five seeds pass all 47; the check that misses most is x2 after the third solve,
    tolerance 0.3).  Required here: at least 93 % of all checks, every single check in at least three of the eight solves, and the
    prior-pinned beacons (0.01) always."""
    checks = [(k, atol) for stage, tol in ATOL_ALL.items() for k, atol in ((( stage, v), a) for v, a in tol.items())]
    ok = np.array([[d[k] < atol for k, atol in checks] for d in dists])
    assert ok.mean() >= 0.93, ok.mean()
    assert ok.sum(0).min() >= 3, [(checks[i][0], int(ok[:, i].sum())) for i in np.flatnonzero(ok.sum(0) < 3)]
    for i, (k, _) in enumerate(checks):
        if k[1] in ("l1", "l2"):
            assert ok[:, i].all(), (k, [d[k] for d in dists])
    return ok


def _pts(belief):
    return belief.to_host().pts if hasattr(belief, "to_host") else belief.pts


SEEDS = (1, 2, 3, 4, 5)


def _check(tut3, outs):
    """outs: one solve per seed.  The reference test is itself one stochastic realisation per run; how the mass of x0 splits between
    its two modes random-walks from sweep to sweep (on the oracle 10 of 12 seeds pass every single check, the misses are x0 / x1 by
    < 0.08 when one mode has died out), so each check must hold for the median over the seeds and, individually, for at least
    four seeds out of five; the prior-pinned landmarks (atol 0.01) must hold in every solve."""
    for stage, tol in ATOL.items():
        for v, atol in tol.items():
            dist = np.array([oracle.mmd(tut3[f"{v}_{stage}"][0], out[stage][v], 0.001) for out in outs])
            assert np.median(dist) < atol and (dist < atol).sum() >= len(outs) - 1, (stage, v, dist, atol)
            if v in ("l1", "l2"):
                assert dist.max() < atol, (stage, v, dist)


def test_graph_matches_the_reference_recipe():
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    fg = solver.generateGraph_Tut3()
    assert list(fg.variables) == ["x0", "l1", "l2", "l3"] and len(fg.factors) == 5
    rho = {tuple(ls): f.rho for ls, f in fg.factors if isinstance(f, solver.Point2Point2Range)}
    assert np.isclose(rho[("x0", "l1")], np.hypot(10, 30)) and np.isclose(rho[("x0", "l3")], np.hypot(80, 40))
    solver.tut3_vehicle_drives(fg, "x0", "x1")
    solver.tut3_vehicle_drives(fg, "x1", "x2")
    # x1 sees l1..l4 (all within 150 m), so does x2; one odometry range each
    assert "l4" in fg.variables and len(fg.factors) == 5 + 2 * (1 + 4)


def _bimodal_x0_and_ring_l3(out):
    """what the tutorial is about: after the first solve x0 sits on the two intersections of the l1 and l2 rings and l3 is a ring"""
    x0 = out[1]["x0"]
    near_truth = np.hypot(*(x0 - [0.0, 0.0]).T) < 12
    near_mirror = np.hypot(*(x0 - [36.0, 12.0]).T) < 12
    return near_truth.mean() > 0.1 and near_mirror.mean() > 0.1 and (near_truth | near_mirror).mean() > 0.8 and out[1]["l3"].std(0).min() > 40


def test_tut3_checkpoints_on_oracle_backend(tut3):
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    from oracle.solver_backend import OracleBackend
    outs = [_solve_two_stages(solver, OracleBackend(), seed)[1] for seed in SEEDS]
    _check(tut3, outs)
    assert sum(_bimodal_x0_and_ring_l3(o) for o in outs) >= len(outs) - 1


def test_tut3_all_five_checkpoints_on_oracle_backend(tut3):
    load_package()
    solver = import_module("caesar_jl_b200.solver")
    from oracle.solver_backend import OracleBackend
    ok = _check_all([_solve_five_stages(solver, OracleBackend(), seed, tut3) for seed in SEEDS_ALL])
    assert ok.all(axis=1).sum() >= 2   # whole runs of the reference test that pass outright


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_tut3_all_five_checkpoints_on_device(tut3, precision):
    build()
    cb = load_package()
    solver = import_module("caesar_jl_b200.solver")
    ctx = cb.Context(0, precision)
    for backend in (solver.DeviceBackend(cb, ctx), solver.DeviceResidentBackend(cb, ctx)):
        ok = _check_all([_solve_five_stages(solver, backend, seed, tut3) for seed in SEEDS_ALL])
        assert ok.all(axis=1).sum() >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_tut3_checkpoints_on_device(tut3, precision):
    build()
    cb = load_package()
    solver = import_module("caesar_jl_b200.solver")
    ctx = cb.Context(0, precision)
    for backend in (solver.DeviceBackend(cb, ctx), solver.DeviceResidentBackend(cb, ctx)):
        outs = [_solve_two_stages(solver, backend, seed)[1] for seed in SEEDS]
        _check(tut3, outs)
        assert sum(_bimodal_x0_and_ring_l3(o) for o in outs) >= len(outs) - 1


@pytest.mark.gpu
def test_point2_convolutions_match_oracle():
    build()
    cb = load_package()
    ctx = cb.Context(0, "fp64")
    rng = np.random.default_rng(2)
    src = rng.normal(size=(77, 2)) * 20
    L = np.zeros((3, 3))
    L[:2, :2] = np.linalg.cholesky(np.array([[2.0, 0.3], [0.3, 1.0]]))
    convs = [dict(kind=cb.FACTOR_PRIOR_POINT2, reverse=False, N=300, mean=[10.0, 30.0, 0.0], chol=L, stream=4, src=None, aux=None),
             dict(kind=cb.FACTOR_POINT2POINT2_RANGE, reverse=False, N=300, mean=[31.6, 0, 0], chol=np.diag([2.0, 0, 0]), stream=9, src=src, aux=None),
             dict(kind=cb.FACTOR_POINT2POINT2_RANGE, reverse=True, N=50, mean=[89.4, 0, 0], chol=np.diag([3.0, 0, 0]), stream=10, src=src, aux=None)]
    got = cb.approxConvBatch(convs, seed=11, ctx=ctx)
    for c, g in zip(convs, got):
        want = oracle.approx_conv(c["kind"], c["reverse"], c["N"], c["mean"], c["chol"], 11, stream=c["stream"], src=c["src"])
        assert g.shape == want.shape == (c["N"], 2)
        assert np.allclose(g, want, rtol=0, atol=1e-10)
    r = np.hypot(*(got[1] - src[np.arange(300) % 77]).T)
    assert abs(r.mean() - 31.6) < 0.5 and abs(r.std() - 2.0) < 0.4
    ang = np.arctan2(*(got[1] - src[np.arange(300) % 77]).T[::-1])
    assert np.histogram(ang, bins=4, range=(-np.pi, np.pi))[0].min() > 45   # uniform direction
