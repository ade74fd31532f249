"""Closed-form / identity checks of the CPU restatement."""
import numpy as np
import pytest

import oracle
from oracle import np_mirror as npm


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, ctr=0 key=0 and ctr=key=ffffffff
    assert [hex(x) for x in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert tuple(oracle.philox([1, 2, 3, 4], [5, 6])) == npm.philox4x32((1, 2, 3, 4), (5, 6))
    for ub in (24, 53):
        assert oracle.uniform(99, 3, 4, 5, 0, ub) == npm.uniform(99, 3, 4, 5, 0, ub)


def test_mmd_identities():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(60, 3)), rng.normal(size=(45, 3)) + 0.5
    assert abs(oracle.mmd(a, a, 0.7)) < 1e-15
    assert np.isclose(oracle.mmd(a, b, 0.7), oracle.mmd(b, a, 0.7), rtol=1e-13)
    assert np.isclose(oracle.mmd(a + 3.0, b + 3.0, 0.7), oracle.mmd(a, b, 0.7), rtol=1e-11)
    assert np.isclose(oracle.mmd(a, b, 0.7), npm.mmd(a, b, 0.7), rtol=1e-12)
    v, s = oracle.mmd(a, b, 0.7, return_sums=True)
    parts = oracle.mmd_rows(a, b, 0.7, (0, 20), (0, 10)) + oracle.mmd_rows(a, b, 0.7, (20, 60), (10, 45))
    assert np.allclose(parts, s, rtol=1e-13)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("backward", [False, True])
def test_se_transform_and_gradient(dim, backward):
    rng = np.random.default_rng(dim)
    nc = 3 if dim == 2 else 6
    c = rng.normal(size=nc) * np.r_[np.ones(dim), 0.4 * np.ones(nc - dim)]
    R, t = oracle.se_matrix(dim, c, backward)
    R2, t2 = npm.se_matrix(dim, c, backward)
    assert np.allclose(R, R2, atol=1e-12) and np.allclose(t, t2, atol=1e-12)
    pts = rng.normal(size=(9, dim))
    assert np.allclose(oracle.transform_points(dim, c, pts, backward), pts @ R.T + t, atol=1e-13)
    # backward undoes forward (ConsolidateRigidTransform.jl:33)
    assert np.allclose(oracle.transform_points(dim, c, oracle.transform_points(dim, c, pts, False), True), pts, atol=1e-12)
    a, b = rng.normal(size=(40, dim)), rng.normal(size=(35, dim))
    vals, grads = oracle.mmd_se_batch(a, b, 0.5, c[None], backward)
    assert np.isclose(vals[0], oracle.mmd(a, oracle.transform_points(dim, c, b, backward), 0.5), rtol=1e-12)
    h = 1e-6
    for k in range(nc):
        e = np.zeros(nc); e[k] = h
        fd = (oracle.mmd_se_batch(a, b, 0.5, (c + e)[None], backward, grad=False)[0]
              - oracle.mmd_se_batch(a, b, 0.5, (c - e)[None], backward, grad=False)[0]) / (2 * h)
        assert np.isclose(grads[0, k], fd, rtol=1e-5, atol=1e-9), (k, grads[0, k], fd)


def test_kde_eval_integrates_to_one_and_matches_mirror():
    rng = np.random.default_rng(1)
    mu, sig = rng.normal(size=(50, 2)), np.array([0.3, 0.5])
    g = np.linspace(-6, 6, 241)
    q = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    p = oracle.kde_eval(mu, sig, q)
    assert abs(p.sum() * (g[1] - g[0]) ** 2 - 1) < 1e-6
    w = rng.uniform(0.1, 1, size=50)
    assert np.allclose(oracle.kde_eval(mu, sig, q[::97], w), npm.kde_eval(mu, sig, q[::97], w), rtol=1e-12)
    assert np.allclose(oracle.kde_eval(mu, sig, q[::97], logp=True), np.log(npm.kde_eval(mu, sig, q[::97])), rtol=1e-10)


def test_loo_bandwidth_is_argmin_of_objective():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(200, 1))
    h = oracle.kde_bw_loo(x)[0]
    hs = np.geomspace(h / 3, h * 3, 61)
    f = [oracle.loo_nll_1d(x[:, 0], v) for v in hs]
    assert abs(hs[int(np.argmin(f))] / h - 1) < 0.05


def test_sample_moments():
    rng = np.random.default_rng(3)
    mu, sig = rng.normal(size=(30, 3)) * [1, 2, 0.3], np.array([0.2, 0.1, 0.05])
    s = oracle.kde_sample(mu, sig, 40000, seed=11)
    assert np.allclose(s.mean(0), mu.mean(0), atol=0.03)
    assert np.allclose(s.var(0), mu.var(0) + sig**2, rtol=0.05)
    sc = oracle.kde_sample(np.array([[3.1]]), np.array([0.2]), 1000, seed=5, circular=[1])
    assert np.abs(sc).max() <= np.pi + 1e-12 and (sc < -2.5).any()


def test_tree_matches_mirror_and_moments():
    rng = np.random.default_rng(4)
    for N in (1, 2, 3, 7, 64, 100):
        pts, bw = rng.normal(size=(N, 3)) * [3, 1, 0.2], np.array([0.1, 0.2, 0.05])
        w = rng.uniform(0.5, 1.5, size=N)
        t, m = oracle.Tree(pts, bw, w), npm.build_tree(pts, bw, w)
        assert t.depth == m["depth"] and (t.perm == m["perm"]).all()
        assert sorted(t.perm) == list(range(N))
        for l in range(t.depth + 1):
            assert t.cnt[l] == min(2**l, N) if l < t.depth else t.cnt[l] == N
            assert np.allclose(t.mean[l], m["mean"][l], rtol=1e-12, atol=1e-14)
            assert np.allclose(t.var[l], m["var"][l], rtol=1e-10, atol=1e-14)
            assert np.isclose(t.wt[l].sum(), 1.0)
        wn = w / w.sum()
        assert np.allclose(t.mean[0][0], wn @ pts, atol=1e-12)
        assert np.allclose(t.var[0][0], wn @ (pts - wn @ pts) ** 2 + bw**2, rtol=1e-10)


def _gauss_kde(rng, m, s, N):
    """single Gaussian N(m, diag s^2) as an N-kernel KDE with a small (Silverman) bandwidth"""
    pts = m + rng.normal(size=(N, len(m))) * s
    bw = 1.06 * pts.std(0) * N ** (-0.2)
    return pts, bw


def exact_product_moments(dens):
    """Exact mean/var and per-density label marginals of the product of D Gaussian-mixture KDEs (enumeration)."""
    D, d = len(dens), dens[0][0].shape[1]
    Ns = [len(x[0]) for x in dens]
    grids = np.meshgrid(*[np.arange(n) for n in Ns], indexing="ij")
    idx = [g.reshape(-1) for g in grids]
    lam = [1.0 / np.asarray(x[1]) ** 2 for x in dens]
    L = sum(lam)
    mu = sum(x[0][i] * l for x, i, l in zip(dens, idx, lam)) / L  # combos x d
    # log weight of a combination: sum_j log N(mu_j; m, s_j^2) integrated over x
    logw = sum((-0.5 * l * (x[0][i] - mu) ** 2).sum(1) for x, i, l in zip(dens, idx, lam))
    w = np.exp(logw - logw.max())
    w /= w.sum()
    mean = w @ mu
    var = w @ (mu - mean) ** 2 + 1.0 / L
    marg = [np.bincount(i, weights=w, minlength=n) for i, n in zip(idx, Ns)]
    return mean, var, marg


def test_product_matches_exact_enumeration():
    """The product of KDEs is a finite Gaussian mixture; for small N its moments and label marginals are
    available exactly.  The multiscale Gibbs sampler (App. A.3) must reproduce them."""
    rng = np.random.default_rng(5)
    ms = [np.array([0.0, 1.0]), np.array([1.0, -0.5]), np.array([0.4, 0.2])]
    ss = [np.array([1.0, 0.5]), np.array([0.7, 0.9]), np.array([1.2, 0.6])]
    for D, N in ((2, 300), (3, 50)):
        dens = [_gauss_kde(rng, m, s, N) for m, s in zip(ms[:D], ss[:D])]
        mean, var, marg = exact_product_moments(dens)
        # Niter=1 (what IIF passes) is a short chain: small residual bias, ~0.1 sigma
        out, labels = oracle.product(dens, 12000, seed=123)
        assert np.allclose(out.mean(0), mean, atol=0.2 * np.sqrt(var).max())
        assert np.allclose(out.var(0), var, rtol=0.2)
        # a longer chain converges to the exact mixture
        out, labels = oracle.product(dens, 12000, seed=123, niter=6)
        assert np.allclose(out.mean(0), mean, atol=4 * np.sqrt(var / 12000).max() + 0.01)
        assert np.allclose(out.var(0), var, rtol=0.08)
        for j in range(D):
            emp = np.bincount(labels[:, j], minlength=N) / 12000
            assert 0.5 * np.abs(emp - marg[j]).sum() < 0.08


def test_product_closed_form_two_gaussians():
    """App. A.3 closed-form check: Gaussian messages -> information-weighted mean, (sum S^-1)^-1 variance
    (each S inflated by the KDE bandwidth).  Finite-N KDEs only approximate their Gaussians: loose bounds."""
    rng = np.random.default_rng(8)
    dens = [_gauss_kde(rng, np.array([0.0, 1.0]), np.array([1.0, 0.5]), 1500),
            _gauss_kde(rng, np.array([1.0, -0.5]), np.array([0.7, 0.9]), 1500)]
    out, _ = oracle.product(dens, 6000, seed=1, niter=3)
    S = [x[0].var(0) + x[1] ** 2 for x in dens]
    lamS = sum(1 / v for v in S)
    assert np.allclose(out.mean(0), sum(x[0].mean(0) / v for x, v in zip(dens, S)) / lamS, atol=0.08)
    assert np.allclose(out.var(0), 1 / lamS, rtol=0.15)


def test_product_bimodal_mass_and_circular():
    rng = np.random.default_rng(6)
    # message 1: two modes at -3 (40%) and +3 (60%); message 2: broad Gaussian -> masses ~ preserved
    N = 500
    comp = rng.uniform(size=N) < 0.4
    p1 = np.where(comp, -3.0, 3.0)[:, None] + 0.3 * rng.normal(size=(N, 1))
    p2 = 4.0 * rng.normal(size=(N, 1))
    dens = [(p1, np.array([0.15])), (p2, np.array([1.0]))]
    out, lab = oracle.product(dens, 3000, seed=9)
    _, _, marg = exact_product_moments(dens)
    exact_mass = marg[0][comp].sum()  # exact mass of the left mode in the product mixture
    assert abs(exact_mass - 0.4) < 0.15
    # Niter=1 mixes slowly between well-separated modes; the reference's own mode-mass test allows +-0.2
    # (test/ICRA2022_tests/nongaussian_mixture.jl:61-63).  A long chain converges to the exact mass.
    assert abs(comp[lab[:, 0]].mean() - exact_mass) < 0.2
    assert abs((out[:, 0] < 0).mean() - exact_mass) < 0.2
    out, lab = oracle.product(dens, 3000, seed=9, niter=16)
    assert abs(comp[lab[:, 0]].mean() - exact_mass) < 0.05
    # circular: two messages straddling the +-pi cut multiply to a mode at the cut, not at 0
    a = oracle.np_mirror.wrap_pi(np.pi - 0.2 + 0.1 * rng.normal(size=(200, 1)))
    b = oracle.np_mirror.wrap_pi(-np.pi + 0.2 + 0.1 * rng.normal(size=(200, 1)))
    out, _ = oracle.product([(a, np.array([0.05])), (b, np.array([0.05]))], 1000, seed=2, circular=[1])
    assert (np.abs(out) > 2.8).mean() > 0.99


def test_product_single_density_and_offsets():
    rng = np.random.default_rng(7)
    pts, bw = rng.normal(size=(37, 2)), np.array([0.2, 0.3])
    out, lab = oracle.product([(pts, bw)], 500, seed=3)
    assert np.abs(np.bincount(lab[:, 0], minlength=37) / 500 - 1 / 37).max() < 0.05
    dens = [(pts, bw), (rng.normal(size=(64, 2)), np.array([0.3, 0.3]))]
    full, lf = oracle.product(dens, 20, seed=3)
    part, lp = oracle.product(dens, 8, seed=3, sample_offset=12)
    assert np.array_equal(full[12:], part) and np.array_equal(lf[12:], lp)


def _room(n, rng):
    """Three orthogonal walls with 1 cm roughness: a cloud with well-defined normals."""
    k = n // 3
    a = np.c_[rng.uniform(0, 4, k), rng.uniform(0, 4, k), 0.01 * rng.normal(size=k)]
    b = np.c_[rng.uniform(0, 4, k), 0.01 * rng.normal(size=k), rng.uniform(0, 4, k)]
    c = np.c_[0.01 * rng.normal(size=n - 2 * k), rng.uniform(0, 4, n - 2 * k), rng.uniform(0, 4, n - 2 * k)]
    return np.r_[a, b, c]


def test_icp_restatement_recovers_a_known_rigid_transform():
    """simpleICP (R:src/3rdParty/_PCL/services/ICP_Simple.jl:239-332) on a structured scene: the moving cloud is
    an independent sampling of the same walls moved by T^-1, so fix_H_mov must come out as T."""
    rng = np.random.default_rng(0)
    fix, src = _room(6000, rng), _room(6000, rng)
    truth = np.array([0.15, -0.1, 0.08, 0.03, -0.02, 0.05])
    mov = oracle.transform_points(3, truth, src, True)
    H, st = oracle.icp_align(fix, mov)
    back = oracle.transform_points(3, truth, np.eye(3), False) - oracle.transform_points(3, truth, np.zeros((1, 3)), False)
    assert np.abs(H[0, :3, :3] - back.T).max() < 2e-3          # rotation
    assert np.abs(H[0, :3, 3] - truth[:3]).max() < 6e-3        # translation
    assert H[0, 3].tolist() == [0, 0, 0, 1] and 2 <= st[0, 3] <= 25 and st[0, 2] > 300 and abs(st[0, 0]) < 1e-3
    # normals of wall points are the wall normals (sign fixed: largest component positive); corners are rejected by planarity
    sel, nrm, pl = oracle.icp_normals(fix, 500, 50)
    flat = pl > 0.6
    big = np.abs(nrm[flat]).argmax(axis=1)
    v = nrm[flat][np.arange(flat.sum()), big]
    assert flat.sum() > 300 and np.median(v) > 0.999 and (v > 0.8).all()   # tilted only next to the wall intersections
    assert sel[0] == 0 and sel[-1] == 5999 and np.all(np.diff(sel) > 0)
    # the perturbed particles of getSample (ScatterAlignPose2.jl:195-196) end at H_k = T o T(dstb_k)
    pert = 0.05 * rng.normal(size=(3, 6))
    Hk, stk = oracle.icp_align(fix, mov, perturb=pert)
    for k in range(3):
        Tp = np.eye(4)
        Tp[:3, :3] = (oracle.transform_points(3, pert[k], np.eye(3), False) - oracle.transform_points(3, pert[k], np.zeros((1, 3)), False)).T
        Tp[:3, 3] = pert[k, :3]
        assert np.abs(Hk[k] - H[0] @ Tp).max() < 1.5e-2, k
    with pytest.raises(ValueError):
        oracle.icp_align(fix, mov, correspondences=5)


def test_point2_convolutions_prior_and_range_ring():
    """PriorPoint2 and Point2Point2Range (the factors of R:test/ICRA2022_tests/insufficient_data.jl:104-134): a Gaussian, and a ring
    of radius ~ N(rho, sigma) around the known point with a uniform direction, the same in both directions of the factor."""
    L = np.zeros((3, 3))
    L[:2, :2] = np.linalg.cholesky(np.array([[2.0, 0.6], [0.6, 1.0]]))
    pr = oracle.approx_conv(5, False, 20000, [10.0, 30.0, 0.0], L, seed=3, stream=1)
    assert pr.shape == (20000, 2)
    assert np.allclose(pr.mean(0), [10.0, 30.0], atol=0.05) and np.allclose(np.cov(pr.T), [[2.0, 0.6], [0.6, 1.0]], atol=0.06)
    src = np.array([[10.0, 30.0], [10.5, 29.0], [9.0, 31.0]])
    for reverse in (False, True):
        ring = oracle.approx_conv(6, reverse, 6000, [31.6, 0, 0], np.diag([2.0, 0, 0]), seed=4, stream=2, src=src)
        rel = ring - src[np.arange(6000) % 3]
        r, th = np.hypot(*rel.T), np.arctan2(rel[:, 1], rel[:, 0])
        assert abs(r.mean() - 31.6) < 0.1 and abs(r.std() - 2.0) < 0.1
        assert np.histogram(th, bins=8, range=(-np.pi, np.pi))[0].min() > 0.8 * 6000 / 8   # uniform direction
    # same counters, same draws: the reverse flag does not change a symmetric factor
    a = oracle.approx_conv(6, False, 50, [5.0, 0, 0], np.diag([0.1, 0, 0]), seed=9, stream=0, src=src)
    b = oracle.approx_conv(6, True, 50, [5.0, 0, 0], np.diag([0.1, 0, 0]), seed=9, stream=0, src=src)
    assert np.array_equal(a, b)


def test_heatmap_grid_density_counterpart_places_mass_where_the_image_has_it():
    x, y = np.linspace(-5, 5, 41), np.linspace(-3, 3, 25)
    gx, gy = np.meshgrid(x, y, indexing="ij")
    img = np.exp(-0.5 * ((gx - 2.0) ** 2 + (gy + 1.0) ** 2) / 0.3) - 0.05   # negative background carries no mass
    pts, bw = oracle.heatmap_grid_density(img, (x, y), 2000, seed=1)
    assert pts.shape == (2000, 2) and bw.shape == (2,) and (bw > 0).all()
    assert np.allclose(pts.mean(0), [2.0, -1.0], atol=0.05)
    assert np.all(np.hypot(pts[:, 0] - 2.0, pts[:, 1] + 1.0) < 2.5)
    # reproducible: the Philox stream is keyed by the seed alone
    p2, _ = oracle.heatmap_grid_density(img, (x, y), 2000, seed=1)
    assert np.array_equal(pts, p2)
