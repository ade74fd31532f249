"""Oracle vs the reference's own stored fixtures."""
import numpy as np

import oracle
from oracle import np_mirror as npm


def test_loo_bandwidth_reproduces_upstream_fixture_bandwidths(tut3):
    """Golden pin: every tut3 belief stores the bw that upstream manikde! chose for its pts.
    Upstream searches with golden-section tol 1e-2, so agreement is expected to <1e-2 relative."""
    worst = 0.0
    for name, (pts, bw) in tut3.items():
        got = oracle.kde_bw_loo(pts)
        worst = max(worst, np.abs(got / bw - 1).max())
    assert worst < 1e-2, worst


def test_mmd_between_stored_stages_within_reference_tolerances(tut3):
    """test/ICRA2022_tests/insufficient_data.jl:143-248: mmd(stored, solved) < atol with default bw 0.001.
    Successive stored solves of the prior-pinned landmarks must satisfy the tightest tolerance (0.01)."""
    for v in ("l1", "l2"):
        for s in range(1, 5):
            a, b = tut3[f"{v}_{s}"][0], tut3[f"{v}_{s + 1}"][0]
            assert oracle.mmd(a, b, 0.001) < 0.01
    # x0 files 1..4 hold identical points -> mmd exactly 0
    assert abs(oracle.mmd(tut3["x0_1"][0], tut3["x0_2"][0], 0.001)) < 1e-15
    # distinct variables are far apart under the same kernel
    assert oracle.mmd(tut3["l1_1"][0], tut3["l2_1"][0], 0.001) > 0.25


def test_c_oracle_matches_numpy_mirror_on_fixtures(tut3):
    a, sa = tut3["l3_2"]
    b, sb = tut3["l3_3"]
    assert np.isclose(oracle.mmd(a, b, 0.001), npm.mmd(a, b, 0.001), rtol=1e-12, atol=1e-15)
    q = np.vstack([a[:7], b[:7]])
    assert np.allclose(oracle.kde_eval(a, sa, q), npm.kde_eval(a, sa, q), rtol=1e-12)
    for h in (0.3, 2.0, 9.0):
        assert np.isclose(oracle.loo_nll_1d(a[:, 0], h), npm.loo_nll_1d(a[:, 0], h), rtol=1e-12)
    assert np.isclose(oracle.kde_bw_loo(a)[1], npm.bw_loo_1d(a[:, 1]), rtol=1e-12)


def test_product_of_fixture_beliefs_matches_mirror_labels(tut3):
    """Same Philox stream -> the C oracle and the pure-Python mirror must draw identical labels."""
    dens = [tut3["l3_2"], tut3["l3_3"], tut3["l3_4"]]
    p1, l1 = oracle.product(dens, 24, seed=7)
    p2, l2 = npm.product(dens, 24, seed=7)
    assert (l1 == l2).all()
    assert np.allclose(p1, p2, rtol=1e-10, atol=1e-10)
