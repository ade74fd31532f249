"""BASELINE.json configs as parity-test cases on the device: C1 heatmap ScatterAlignPose2, C3 ScatterAlignPose3 clouds
(mmd + alignment), C4 multihypo four-mode products (N=1000)."""
import numpy as np
import pytest

import oracle
from __graft_entry__ import build, load_package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    build()
    return load_package()


@pytest.fixture(scope="module", params=["fp32", "fp64"])
def ctx(request, cb):
    return cb.Context(0, request.param)


@pytest.fixture(scope="module")
def ctx32(cb):
    return cb.Context(0, "fp32")


def test_c1_heatmap_scatter_align_pose2(cb, ctx):
    """R:test/testScatterAlignPose2.jl:19-69: 301x301 heatmaps of three Gaussians, true pCq = [5, 0, pi/8],
    N=1000, sample_count=100, bw=1.0; recovery within 1.5 m / 0.2 rad (:68-69)."""
    x = np.arange(-15, 15.0001, 0.1)
    gx, gy = np.meshgrid(x, x, indexing="ij")

    def g(px, py):
        f = lambda cx, cy, s: np.exp(-0.5 * ((px - cx) ** 2 + (py - cy) ** 2) / s) / (2 * np.pi * s)
        return f(3.0, 0.0, 0.1) + f(8.0, 0.0, 0.4) + f(0.0, 5.0, 0.1)

    pCq = np.array([5.0, 0.0, np.pi / 8])
    R, t = oracle.se_matrix(2, pCq, backward=True)  # qTp = inv(exp(pCq))
    im1 = g(gx, gy)
    im2 = g(R[0, 0] * gx + R[0, 1] * gy + t[0], R[1, 0] * gx + R[1, 1] * gy + t[1])
    sap = cb.ScatterAlignPose2_from_images(im1, im2, (x, x), sample_count=100, bw=1.0, N=1000, ctx=ctx)
    assert sap.align.cloud1.N == 1000 and sap.align.cloud1.d == 2
    s, = cb.sample(sap.align.cloud1, 10, ctx=ctx)
    assert s.shape == (10, 2)  # `sample(sap.align.cloud2,10)[1] isa AbstractArray` (:56-57)
    cf = cb.CalcFactor(sap, cb.preambleCache(None, [], sap))
    rng = np.random.default_rng(0)
    x0 = pCq + np.c_[0.7 * rng.normal(size=(24, 2)), 0.1 * rng.normal(size=24)]
    coords, score, iters = cb.getSample(cf, x0=x0, seed=3, ctx=ctx, return_info=True)
    best = coords[np.argmin(score)]
    assert np.abs(best[:2] - pCq[:2]).max() < 1.5 and abs(best[2] - pCq[2]) < 0.2
    assert np.abs(np.median(coords, axis=0) - pCq).max() < 0.5


def test_heatmap_grid_density_matches_oracle_counterpart(cb):
    """the heatmap construction (cell weighting, weighted device sampling with the block-scan CDF, LOO
    bandwidth) against its CPU counterpart on the C1 grid (301 x 301 cells, N = 1000) and on a ragged small one."""
    ctx64 = cb.Context(0, "fp64")
    x = np.arange(-15, 15.0001, 0.1)
    gx, gy = np.meshgrid(x, x, indexing="ij")
    f = lambda cx, cy, s: np.exp(-0.5 * ((gx - cx) ** 2 + (gy - cy) ** 2) / s) / (2 * np.pi * s)
    im = f(3.0, 0.0, 0.1) + f(8.0, 0.0, 0.4) + f(0.0, 5.0, 0.1)
    hgd = cb.HeatmapGridDensity(im, (x, x), N=1000, seed=5, ctx=ctx64)
    want_pts, want_bw = oracle.heatmap_grid_density(im, (x, x), 1000, 5)
    assert hgd.densityFnc.pts.shape == (1000, 2)
    assert np.allclose(hgd.densityFnc.pts, want_pts, rtol=0, atol=1e-9)
    assert np.allclose(hgd.densityFnc.bw, want_bw, rtol=1e-6)
    # mass lands on the three blobs in proportion to their integrals (1 : 1 : 1)
    share = [np.mean(np.hypot(*(want_pts - c).T) < 2.5) for c in ((3.0, 0.0), (8.0, 0.0), (0.0, 5.0))]
    assert np.allclose(share, 1 / 3, atol=0.06)
    rng = np.random.default_rng(0)
    xs, ys = np.linspace(0, 1, 7), np.linspace(-2, 2, 13)
    small = rng.uniform(-0.2, 1.0, size=(7, 13))   # negative cells carry no mass
    h2 = cb.HeatmapGridDensity(small, (xs, ys), N=257, seed=9, ctx=ctx64)
    p2, b2 = oracle.heatmap_grid_density(small, (xs, ys), 257, 9)
    assert np.allclose(h2.densityFnc.pts, p2, rtol=0, atol=1e-9) and np.allclose(h2.densityFnc.bw, b2, rtol=1e-6)


def test_c3_scatter_align_pose3_clouds(cb, ctx):
    """R:test/testScatterAlignPose3.jl:29-30 / testScatterAlignPose2.jl:271-299: randn(P,3) clouds, moved cloud offset
    pCq = [2,0,2,0,0,pi/10], kde bandwidth 0.1 per axis (testScatterAlignParched.jl:41-42), mmd bw = 1."""
    rng = np.random.default_rng(1)
    P = 10000
    X = rng.normal(size=(P, 3)) * [3.0, 1.5, 0.7]
    pCq = np.array([2.0, 0, 2.0, 0, 0, np.pi / 10])
    Xm = oracle.transform_points(3, pCq, X + 0.02 * rng.normal(size=X.shape), backward=False)
    # full mmd on the 1e4-point clouds against the oracle (3e8 pairs: seconds on the CPU)
    want, sums = oracle.mmd(X, Xm, 1.0, return_sums=True)
    got = cb.mmd_sums(X, Xm, 1.0, ctx=ctx)
    assert np.allclose(got, sums, rtol=1e-5 if ctx.precision == "fp32" else 1e-12)
    # alignment from sampled subsets (sample_count = 500)
    c1 = cb.MKD("Point3", X, [0.1, 0.1, 0.1])
    c2 = cb.MKD("Point3", Xm, [0.1, 0.1, 0.1])
    sap = cb.ScatterAlignPose3(cloud1=c1, cloud2=c2, sample_count=500, bw=1.0)
    cf = cb.CalcFactor(sap, cb.preambleCache(None, [], sap))
    x0 = pCq + np.c_[0.3 * rng.normal(size=(8, 3)), 0.05 * rng.normal(size=(8, 3))]
    coords, score, iters = cb.getSample(cf, x0=x0, seed=4, ctx=ctx, return_info=True)
    est = np.median(coords, axis=0)
    assert np.abs(est[:3] - pCq[:3]).max() < 0.3 and np.abs(est[3:] - pCq[3:]).max() < 0.1
    # 12-element SE(3) point from the coordinates (reference: the Pose3 sample is a 12-element ArrayPartition, :314-317)
    R, t = oracle.se_matrix(3, est, backward=False)
    assert R.size + t.size == 12 and np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def _four_mode_message(rng, N, centres, spread):
    comp = rng.integers(0, len(centres), size=N)
    pts = centres[comp] + rng.normal(size=(N, 3)) * spread
    pts[:, 2] = (pts[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return pts, 1.06 * pts.std(0) * N ** -0.2 * 0.5


def test_c4_multihypo_four_mode_products(cb, ctx):
    """Standalone C4 product: D=3 messages, each a 4-component mixture in (x, y, theta), N=1000;
    only one corner is common to all three, so the product must concentrate there.  Device vs oracle: mode masses."""
    rng = np.random.default_rng(2)
    corners = np.array([[0, 0, 0.0], [10, 0, 1.5], [10, 10, 3.0], [0, 10, -1.5]])
    spread = np.array([0.4, 0.4, 0.1])
    msgs = []
    for j in range(3):
        cs = corners.copy()
        cs[1:] += rng.normal(size=(3, 3)) * [3.0, 3.0, 0.5] * (j + 1)  # the other hypotheses disagree between messages
        msgs.append(_four_mode_message(rng, 1000, cs, spread))
    N = 1000
    near = lambda p: (np.linalg.norm(p[:, :2] - corners[0, :2], axis=1) < 2.0).mean()
    ub = 53 if ctx.precision == "fp64" else 24
    # default levelDown rule (weight-proportional random child): the chains start each level in unrelated branches of the three
    # trees and, with Niter = 1, only part of them has found the common corner by the leaf level -- device and oracle alike
    pts_r, lab_r = cb.product_points(msgs, (0, 0, 1), N, seed=6, ctx=ctx)
    opts_r, olab_r = oracle.product(msgs, N, seed=6, circular=[0, 0, 1], ubits=ub)
    assert abs(near(pts_r) - near(opts_r)) < 0.05 and near(pts_r) > 0.3
    assert (lab_r == olab_r).all(axis=1).mean() > (0.99 if ctx.precision == "fp64" else 0.85)
    # first-child rule (the common corner happens to sit in the first branches): everything ends there
    ctx.set_level_down(0)
    oracle.set_level_down(0)
    try:
        pts, lab = cb.product_points(msgs, (0, 0, 1), N, seed=6, ctx=ctx)
        opts, olab = oracle.product(msgs, N, seed=6, circular=[0, 0, 1], ubits=ub)
    finally:
        ctx.set_level_down(1)
        oracle.set_level_down(1)
    assert near(pts) > 0.9 and abs(near(pts) - near(opts)) < 0.05
    assert (lab == olab).all(axis=1).mean() > (0.99 if ctx.precision == "fp64" else 0.85)
    # belief-level agreement under the reference's mmd metric (multihypo_corners.jl:201-202 uses < 0.25)
    assert cb.mmd(pts, opts, bw=0.001, ctx=ctx) < 0.01
    # density thresholds at the mode / away from it (multihypo_corners.jl:66-90: > 0.15 at modes, < 0.02 elsewhere)
    m = cb.manikde("Pose2", pts, ctx=ctx)
    # evaluated on the position marginal: rebuild a Point2 belief from the first two coordinates
    m2 = cb.manikde("Point2", pts[:, :2], ctx=ctx)
    assert m2(np.array([[0.0, 0.0]]), ctx=ctx)[0] > 0.15 and m2(np.array([[10.0, 10.0]]), ctx=ctx)[0] < 0.02
    assert m.d == 3 and (m.bw > 0).all()


def test_partial_dimension_product(cb, ctx):
    """AMP multiplies inputs that inform only some coordinates dimension-wise: a full Pose2
    belief times a partial prior on (x, y) only must tighten x, y and leave theta's spread untouched."""
    rng = np.random.default_rng(7)
    full = cb.manikde("Pose2", np.c_[rng.normal(size=(400, 2)) * 2.0, 0.3 * rng.normal(size=400)], ctx=ctx)
    prior_pts = np.c_[1.0 + 0.5 * rng.normal(size=(400, 2)), np.zeros(400)]
    prior = cb.ManifoldKernelDensity("Pose2", prior_pts, [0.2, 0.2, 1.0], partial=(1, 2))
    p = cb.manifoldProduct([full, prior], N=400, seed=3, ctx=ctx)
    assert p.pts.shape == (400, 3) and (p.bw > 0).all()
    # x, y: product of N(0, 2^2) and N(1, 0.5^2) -> mean ~ 0.94, std ~ 0.49 (plus KDE smoothing)
    assert np.abs(p.pts[:, :2].mean(0) - 0.94).max() < 0.2 and np.abs(p.pts[:, :2].std(0) - 0.5).max() < 0.15
    # theta: only `full` informs it -> its samples are passed through
    assert abs(p.pts[:, 2].std() - full.pts[:, 2].std()) < 0.05
    o, _ = oracle.product([(full.pts[:, :1], full.bw[:1]), (prior.pts[:, :1], prior.bw[:1])], 400, seed=3, stream=0,
                          ubits=53 if ctx.precision == "fp64" else 24)
    assert abs(o.mean() - p.pts[:, 0].mean()) < 0.1


def test_async_ticket_matches_synchronous_batch(cb, ctx):
    rng = np.random.default_rng(8)
    groups = [[cb.MKD("Point2", rng.normal(size=(120, 2)) + j, np.array([0.3, 0.3])) for j in range(D)] for D in (2, 3)]
    sync = cb.manifoldProducts(groups, N=120, seed=11, streams=[5, 6], ctx=ctx)
    pend = cb.manifoldProducts_submit(groups, N=120, seed=11, streams=[5, 6], ctx=ctx)
    other = cb.mmd(groups[0][0], groups[0][1], ctx=ctx)  # the context stays usable while a ticket is outstanding
    res = pend.wait()
    for a, b in zip(sync, res):
        assert np.array_equal(a.pts, b.pts) and np.allclose(a.bw, b.bw, rtol=1e-12)
    assert np.isfinite(other)
    with pytest.raises(cb.CB2Error):
        pend.wait()  # a ticket can be waited on once


def test_zmq_multiply_distributions_verb(cb, ctx):
    """The request example shipped with the reference (R:ext/services/additional.jl:34)."""
    import json
    from importlib import import_module
    t2 = ('{"payload":{"weights":[{"distType":"BallTreeDensity","kernelType":"Gaussian","weights":[1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0],'
          '"bandwidth":[0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0],"points":[-0.8272759752535751,1.90340735325013,0.5617238720356695,'
          '-1.2610130840544629,-0.5196580171571705,0.47808234574398156,-0.5034354386657888,-0.1289061606999511,-0.6283440136328498,'
          '-0.04520468574277743],"dim":1},{"distType":"BallTreeDensity","kernelType":"Gaussian","weights":[1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0,1.0],'
          '"bandwidth":[0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0],"points":[0.6986754683541115,0.4206784705438769,-0.01798641761511103,'
          '0.2862258908507545,0.5021835960347083,-1.4424531814460273,0.7667313775874391,0.13048774810077268,-0.004778323705883176,'
          '-0.6408606329686699],"dim":1}]},"request":"multiplyDistributions"}')
    ser = import_module("caesar_jl_b200.serialization")
    reply = ser.multiplyDistributions(json.loads(t2), seed=1, ctx=ctx)
    assert reply["status"] == "OK"
    p = reply["payload"]
    assert p["dim"] == 1 and p["kernelType"] == "Gaussian" and len(p["points"]) == 10 and len(p["weights"]) == 10
    assert abs(sum(p["weights"]) - 1.0) < 1e-12 and p["bandwidth"][0] > 0 and np.isfinite(p["points"]).all()
    json.dumps(reply)  # serialisable as is


def test_approx_conv_closed_form_factors_match_oracle(cb, ctx):
    """per-particle factor solves for Pose2 prior, Pose2Pose2 (both directions) and
    Pose2Point2BearingRange (both directions) -- same Philox draws on the device and in the oracle."""
    rng = np.random.default_rng(12)
    poses = np.c_[rng.normal(size=(100, 2)) * 3, rng.uniform(-np.pi, np.pi, size=100)]
    lms = rng.normal(size=(57, 2)) * 5
    L3 = np.linalg.cholesky(np.diag([0.1, 0.2, 0.05]) ** 2 + 0.001)
    convs = [
        dict(kind=cb.FACTOR_PRIOR_POSE2, reverse=False, N=300, mean=[1.0, -2.0, 3.1], chol=L3, stream=1, src=None, aux=None),
        dict(kind=cb.FACTOR_POSE2POSE2, reverse=False, N=100, mean=[10.0, 0, np.pi / 3], chol=L3, stream=2, src=poses, aux=None),
        dict(kind=cb.FACTOR_POSE2POSE2, reverse=True, N=250, mean=[10.0, 0, np.pi / 3], chol=L3, stream=3, src=poses, aux=None),
        dict(kind=cb.FACTOR_POSE2POINT2_BEARINGRANGE, reverse=False, N=100, mean=[0.3, 20.0, 0], chol=np.diag([0.1, 1.0, 0]), stream=4,
             src=poses, aux=None),
        dict(kind=cb.FACTOR_POSE2POINT2_BEARINGRANGE, reverse=True, N=128, mean=[0.3, 20.0, 0], chol=np.diag([0.1, 1.0, 0]), stream=5,
             src=lms, aux=poses),
    ]
    outs = cb.approxConvBatch(convs, seed=99, ctx=ctx)
    for c, o in zip(convs, outs):
        want = oracle.approx_conv(c["kind"], c["reverse"], c["N"], c["mean"], c["chol"], 99, stream=c["stream"], src=c["src"], aux=c["aux"])
        assert o.shape == want.shape and np.allclose(o, want, atol=1e-12)
    # closed-form sanity: a forward then a reverse convolution with zero noise returns the source pose
    z = np.zeros((3, 3))
    fwd = cb.approxConvBatch([dict(kind=1, reverse=False, N=100, mean=[10.0, 0, np.pi / 3], chol=z, stream=0, src=poses, aux=None)], ctx=ctx)[0]
    back = cb.approxConvBatch([dict(kind=1, reverse=True, N=100, mean=[10.0, 0, np.pi / 3], chol=z, stream=0, src=fwd, aux=None)], ctx=ctx)[0]
    d = back - poses
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() < 1e-12


def test_mmd_between_pose_beliefs_uses_manifold_distance(cb, ctx):
    """mmd(X2_a, belief) on Pose2 beliefs (R:test/ICRA2022_tests/multihypo_corners.jl:201-202): wrapped angular
    differences with the product-metric weight; identical beliefs across the +-pi cut must be at distance ~0."""
    rng = np.random.default_rng(21)
    a = np.c_[rng.normal(size=(300, 2)), np.pi - 0.05 + 0.1 * rng.normal(size=300)]
    a[:, 2] = (a[:, 2] + np.pi) % (2 * np.pi) - np.pi           # straddles the cut
    b = a + np.c_[0.01 * rng.normal(size=(300, 2)), 0.01 * rng.normal(size=300)]
    b[:, 2] = (b[:, 2] + np.pi) % (2 * np.pi) - np.pi
    tol = 1e-5 if ctx.precision == "fp32" else 1e-12
    for bw, circ, d in ((0.5, [0, 0, 1], 3), (0.001, [0, 0, 1], 3), (0.3, [0, 0, 0, 1, 1, 1], 6)):
        A = a if d == 3 else np.c_[a, a[:, ::-1]]
        B = b if d == 3 else np.c_[b, b[:, ::-1]]
        A[:, 3:] = (A[:, 3:] + np.pi) % (2 * np.pi) - np.pi
        B[:, 3:] = (B[:, 3:] + np.pi) % (2 * np.pi) - np.pi
        want = oracle.mmd_manifold(A, B, circ, bw)
        got = cb.mmd_manifold(A, B, circ, bw, ctx=ctx)
        assert abs(got - want) <= tol * 4, (bw, d, got, want)
    # two headings 0.04 rad apart on either side of the cut: close on the manifold, 2 pi apart in flat coordinates
    xy = rng.normal(size=(300, 2))
    pa = np.c_[xy, np.pi - 0.02 + 0.005 * rng.normal(size=300)]
    pb = np.c_[xy, -np.pi + 0.02 + 0.005 * rng.normal(size=300)]
    A, B = cb.MKD("Pose2", pa, [0.1, 0.1, 0.05]), cb.MKD("Pose2", pb, [0.1, 0.1, 0.05])
    assert cb.mmd(A, B, bw=0.5, ctx=ctx) < 5e-3
    assert cb.mmd(pa, pb, bw=0.5, ctx=ctx) > 0.1


def _room(n, rng):
    k = n // 3
    a = np.c_[rng.uniform(0, 4, k), rng.uniform(0, 4, k), 0.01 * rng.normal(size=k)]
    b = np.c_[rng.uniform(0, 4, k), 0.01 * rng.normal(size=k), rng.uniform(0, 4, k)]
    c = np.c_[0.01 * rng.normal(size=n - 2 * k), rng.uniform(0, 4, n - 2 * k), rng.uniform(0, 4, n - 2 * k)]
    return np.r_[a, b, c]


@pytest.mark.parametrize("shape", ["randn", "room", "small"])
def test_icp_batch_matches_the_cpu_restatement(cb, ctx32, shape):
    """`_PCL.alignICP_Simple` for K perturbed particles as one device batch == oracle/cb2_oracle.c (FP64 on both
    sides; sums are taken in a different order, hence the 1e-9).  'randn' is the reference's own test input
    (R:test/testScatterAlignPose3.jl:29-30: X_fix = randn(n, 3), X_mov = randn(n, 3); it only has to run), 'room' a structured
    scene with a known answer, 'small' has fewer points than correspondences / an odd count for the median."""
    rng = np.random.default_rng(31)
    if shape == "randn":
        fix, mov, kw = rng.normal(size=(3000, 3)), rng.normal(size=(3000, 3)), {}
    elif shape == "room":
        fix, kw = _room(6000, rng), {}
        mov = oracle.transform_points(3, np.array([0.15, -0.1, 0.08, 0.03, -0.02, 0.05]), _room(5000, rng), True)
    else:
        fix, mov, kw = _room(333, rng), _room(400, rng) + 0.02, dict(correspondences=500, neighbors=12, max_iterations=7)
    pert = np.r_[np.zeros((1, 6)), 0.2 * rng.normal(size=(4, 6)) * (0.25 if shape == "room" else 1.0)]
    H, st = cb.icp_align_batch(fix, mov, pert, ctx=ctx32, **kw)
    Ho, sto = oracle.icp_align(fix, mov, perturb=pert, **kw)
    assert np.array_equal(st[:, 2:], sto[:, 2:]), (st, sto)            # same correspondences kept, same iteration count
    assert np.allclose(H, Ho, rtol=0, atol=1e-9) and np.allclose(st[:, :2], sto[:, :2], rtol=1e-7, atol=1e-12)
    assert np.allclose(H[:, 3], [0, 0, 0, 1]) and np.allclose(np.linalg.det(H[:, :3, :3]), 1.0, atol=1e-9)
    if shape == "room":
        assert np.abs(H[0, :3, 3] - [0.15, -0.1, 0.08]).max() < 6e-3


def test_scatter_align_pose3_getsample_icp_branch(cb, ctx32):
    """getSample(cf::CalcFactor{<:ScatterAlignPose3}) with sample_count = -1 (R:ext/Images/ScatterAlignPose2.jl:187-214,
    R:test/testScatterAlignPose3.jl:50-62): K particles, each aligns its own perturbed copy of cloud2."""
    rng = np.random.default_rng(5)
    fix = _room(6000, rng)
    truth = np.array([0.15, -0.1, 0.08, 0.03, -0.02, 0.05])
    mov = oracle.transform_points(3, truth, _room(6000, rng), True)
    sap = cb.ScatterAlignPose3(cloud1=cb.ManifoldKernelDensity("Point3", fix, [0.1, 0.1, 0.1]),
                               cloud2=cb.ManifoldKernelDensity("Point3", mov, [0.1, 0.1, 0.1]), sample_count=-1)
    cf = cb.CalcFactor(sap, cb.preambleCache(None, [], sap))
    coords, score, iters = cb.getSample(cf, K=16, seed=3, ctx=ctx32, return_info=True)
    assert coords.shape == (16, 6) and (iters > 0).all() and np.abs(score).max() < 1e-2
    # every particle aligned its own bumped copy: undoing the bump (first order) gives the true transform
    dstb = 0.2 * np.random.default_rng(3).standard_normal((16, 6))
    assert np.abs(np.median(coords - dstb, axis=0) - truth).max() < 0.05
    with pytest.raises(cb.CB2Error):
        cb.icp_align_batch(fix, mov, correspondences=5, ctx=ctx32)
    with pytest.raises(ValueError):
        cb.icp_align_batch(fix[:, :2], mov, ctx=ctx32)


# ------------------------------------------------------------------------------------------------ round 2: headline sizes
def test_mmd_tile_shards_add_up_to_the_full_sums(cb, ctx):
    """Multi-GPU mmd keeps the symmetry of the self terms: the tile list is dealt round-robin (cb2_mmd_sums_shard); the shards
    of any world size add up to the full sums, and one shard of one is the plain call."""
    rng = np.random.default_rng(7)
    a, b = rng.normal(size=(5000, 3)), rng.normal(size=(3100, 3)) + 0.5
    full = cb.mmd_sums(a, b, 0.8, ctx=ctx)
    want = oracle.mmd(a, b, 0.8, return_sums=True)[1]
    tol = 1e-5 if ctx.precision == "fp32" else 1e-12
    assert np.allclose(full, want, rtol=tol)
    assert np.allclose(cb.mmd_sums_shard(a, b, 0.8, 0, 1, ctx=ctx), full, rtol=1e-14)
    for world in (2, 3, 8):
        parts = [cb.mmd_sums_shard(a, b, 0.8, r, world, ctx=ctx) for r in range(world)]
        assert np.allclose(np.sum(parts, axis=0), full, rtol=1e-6 if ctx.precision == "fp32" else 1e-13)
    with pytest.raises(cb.CB2Error):
        cb.mmd_sums_shard(a, b, 0.8, 2, 2, ctx=ctx)


def test_mmd_1e5_points_against_the_full_oracle(cb, ctx32):
    """C3 at the headline size: mmd of two 1e5-point ScatterAlignPose3 clouds (3e10 kernel evaluations) against the FULL CPU oracle
    (about 20 s on 16 cores), 1e-5 relative on the three pair sums; and the K2 cost + gradient of one transform at that size."""
    oracle.use_all_cores()
    rng = np.random.default_rng(3)
    P = 100000
    a = rng.normal(size=(P, 3))
    b = oracle.transform_points(3, np.array([2.0, 0, 2.0, 0, 0, np.pi / 10]), rng.normal(size=(P, 3)), backward=False)
    want, sums = oracle.mmd(a, b, 1.0, return_sums=True)
    got_sums = cb.mmd_sums(a, b, 1.0, ctx=ctx32)
    assert np.allclose(got_sums, sums, rtol=1e-5, atol=0), (got_sums, sums)
    scale = (sums[0] + 2 * sums[1] + sums[2]) / P**2
    assert abs(cb.mmd(a, b, 1.0, ctx=ctx32) - want) <= 1e-5 * scale
    parts = [cb.mmd_sums_shard(a, b, 1.0, r, 8, ctx=ctx32) for r in range(8)]
    assert np.allclose(np.sum(parts, axis=0), sums, rtol=1e-5)
    c = np.array([[2.05, 0.02, 1.9, 0.01, -0.02, np.pi / 10 + 0.03]])
    v, g = cb.mmd_se_batch(a, b, 1.0, c, backward=True, ctx=ctx32)
    ov, og = oracle.mmd_se_batch(a, b, 1.0, c, backward=True)
    assert abs(v[0] - ov[0]) <= 1e-5 * scale
    assert np.allclose(g, og, rtol=2e-4, atol=1e-5 * np.abs(og).max())


def test_full_size_properties_mmd_1e5(cb, ctx32):
    """Size-independent properties at the headline size (two 1e5-point clouds, FP32): mmd(a, a) = 0, mmd(a, b) = mmd(b, a), a rigid
    motion applied to both clouds leaves the value unchanged, and the eight tile shards of a multi-GPU split add up to the whole."""
    rng = np.random.default_rng(11)
    P = 100000
    a = rng.normal(size=(P, 3))
    b = rng.normal(size=(P, 3)) * 1.1 + [0.3, 0.0, -0.2]
    s_ab = cb.mmd_sums(a, b, 1.0, ctx=ctx32)
    scale = (s_ab[0] + 2 * s_ab[1] + s_ab[2]) / P**2
    assert abs(cb.mmd(a, a, 1.0, ctx=ctx32)) <= 2e-6 * scale
    v_ab, v_ba = cb.mmd(a, b, 1.0, ctx=ctx32), cb.mmd(b, a, 1.0, ctx=ctx32)
    assert abs(v_ab - v_ba) <= 2e-6 * scale and v_ab > 0
    c = np.array([4.0, -3.0, 2.0, 0.3, -0.2, 0.5])
    ta, tb = oracle.transform_points(3, c, a, backward=False), oracle.transform_points(3, c, b, backward=False)
    assert abs(cb.mmd(ta, tb, 1.0, ctx=ctx32) - v_ab) <= 1e-5 * scale
    parts = np.sum([cb.mmd_sums_shard(a, b, 1.0, r, 8, ctx=ctx32) for r in range(8)], axis=0)
    assert np.allclose(parts, s_ab, rtol=1e-6)


def test_full_size_product_closed_form_gaussians(cb, ctx32):
    """C5 shape (D = 8 messages x N = 4096 kernels, d = 6, N_out = 4096, FP32, tensor-core leaf pass) on inputs with a closed form:
    message j holds samples of N(m_j, S_j), so the product is Gaussian with precision sum_j (S_j + bw_j^2)^-1."""
    rng = np.random.default_rng(12)
    D, N, d = 8, 4096, 6
    dens, prec, pm = [], np.zeros(d), np.zeros(d)
    for j in range(D):
        mj = rng.normal(size=d) * 0.2
        sj = rng.uniform(0.4, 0.8, size=d)
        pts = mj + rng.normal(size=(N, d)) * sj
        bw = 1.06 * pts.std(0) * N ** -0.2
        dens.append((pts, bw))
        lam = 1.0 / (pts.var(0) + bw**2)
        prec += lam
        pm += lam * pts.mean(0)
    want_mean, want_var = pm / prec, 1.0 / prec
    pts, lab = cb.product_points(dens, (0,) * d, 4096, seed=3, ctx=ctx32)
    # (the multiscale sampler with Niter = 1 carries a residual bias of 0.1 - 0.2 sigma on the oracle as well: DESIGN.md sec. 2)
    assert np.all(np.abs(pts.mean(0) - want_mean) < 0.25 * np.sqrt(want_var)), (pts.mean(0), want_mean)
    opts, olab = oracle.product(dens, 512, seed=3, circular=[0] * d, ubits=24)
    assert np.all(np.abs(pts[:512].mean(0) - opts.mean(0)) < 0.1 * np.sqrt(want_var))   # and agrees with the oracle chains
    # variance: with Niter = 1 and eight messages the multiscale sampler is over-dispersed against the exact product (by ~40 % here,
    # on the oracle and the device alike: two sweeps per level do not mix eight label chains fully); the device must show the
    # oracle's spread, and both stay within a factor of the closed form
    assert np.allclose(pts.var(0), opts.var(0), rtol=0.2)
    assert np.all(pts.var(0) > 0.9 * want_var) and np.all(pts.var(0) < 2.0 * want_var)
    # the chains spread over the kernels near the product's mode exactly as the oracle's do (FP32 flips a fraction of a percent)
    assert lab.min() >= 0 and lab.max() < N
    assert np.mean(lab[:512] == olab) > 0.98
    assert len(np.unique(lab[:, 0])) >= len(np.unique(olab[:, 0]))
