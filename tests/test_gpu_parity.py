"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): deterministic kernels 1e-5 relative in FP32 mode, 1e-12 in FP64 mode;
stochastic products: same Philox stream -> (near-)identical label draws in FP64 mode, distribution-level
agreement (moments + fixed-seed MMD two-sample statistic) in FP32 mode."""
import numpy as np
import pytest

import oracle
from __graft_entry__ import build, load_package

pytestmark = pytest.mark.gpu

RTOL = {"fp32": 1e-5, "fp64": 1e-12}


@pytest.fixture(scope="module")
def cb():
    build()
    return load_package()


@pytest.fixture(scope="module", params=["fp32", "fp64"])
def ctx(request, cb):
    return cb.Context(0, request.param)


@pytest.fixture(scope="module")
def ctx64(cb):
    return cb.Context(0, "fp64")


@pytest.fixture(scope="module")
def ctx32(cb):
    return cb.Context(0, "fp32")


# ------------------------------------------------------------------------------------------------ mmd
@pytest.mark.parametrize("d,N,M", [(1, 1, 1), (2, 100, 100), (3, 150, 77), (3, 1025, 4097), (6, 300, 513), (4, 64, 65)])
def test_mmd_matches_oracle(cb, ctx, d, N, M):
    rng = np.random.default_rng(d * 1000 + N)
    a, b = rng.normal(size=(N, d)), rng.normal(size=(M, d)) * 1.2 + 0.4
    for bw in (0.05, 1.0):
        want, sums = oracle.mmd(a, b, bw, return_sums=True)
        got_sums = cb.mmd_sums(a, b, bw, ctx=ctx)
        assert np.allclose(got_sums, sums, rtol=RTOL[ctx.precision], atol=0)
        got = cb.mmd(a, b, bw, ctx=ctx)
        # the three terms nearly cancel: bound the error by the size of the terms
        scale = sums[0] / N**2 + 2 * sums[1] / (N * M) + sums[2] / M**2
        assert abs(got - want) <= RTOL[ctx.precision] * scale


def test_mmd_far_from_origin_and_fixtures(cb, ctx, tut3):
    """Clouds 1e4 m from the origin (FP32 would lose the differences without centring) and the reference's beliefs."""
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(400, 2)) + 1e4, rng.normal(size=(300, 2)) + 1e4 + 0.5
    want, sums = oracle.mmd(a, b, 0.7, return_sums=True)
    assert np.allclose(cb.mmd_sums(a, b, 0.7, ctx=ctx), sums, rtol=RTOL[ctx.precision])
    for x, y in (("l3_2", "l3_3"), ("x1_2", "x1_3"), ("l1_1", "l2_1")):
        A, B = cb.unpackDistribution(cb.packDistribution(cb.MKD("Point2", *tut3[x]))), cb.MKD("Point2", *tut3[y])
        want = oracle.mmd(tut3[x][0], tut3[y][0], 0.001)
        assert abs(cb.mmd(A, B, ctx=ctx) - want) <= RTOL[ctx.precision] * 4
    # reference tolerances (insufficient_data.jl:143-248) hold on the device path too
    assert cb.mmd(cb.MKD("Point2", *tut3["l1_1"]), cb.MKD("Point2", *tut3["l1_2"]), ctx=ctx) < 0.01
    assert cb.isapprox(cb.MKD("Point2", *tut3["x0_1"]), cb.MKD("Point2", *tut3["x0_2"]), ctx=ctx)


def test_mmd_row_shards_add_up(cb, ctx):
    rng = np.random.default_rng(5)
    a, b = rng.normal(size=(3000, 3)), rng.normal(size=(2500, 3))
    full = cb.mmd_sums(a, b, 0.3, ctx=ctx)
    parts = sum(cb.mmd_sums(a, b, 0.3, (r0, r1), (q0, q1), ctx=ctx)
                for (r0, r1), (q0, q1) in zip([(0, 1000), (1000, 1001), (1001, 3000)], [(0, 7), (7, 2000), (2000, 2500)]))
    assert np.allclose(parts, full, rtol=1e-12 if ctx.precision == "fp64" else 1e-6)
    assert np.allclose(parts, oracle.mmd(a, b, 0.3, return_sums=True)[1], rtol=RTOL[ctx.precision])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("backward", [True, False])
def test_mmd_se_batch_value_and_gradient(cb, ctx, dim, backward):
    rng = np.random.default_rng(dim + 10 * backward)
    nc = 3 if dim == 2 else 6
    a = rng.normal(size=(700, dim)) * 2 + 5.0
    b = rng.normal(size=(650, dim)) * 2 + 4.0
    coords = rng.normal(size=(13, nc)) * np.r_[np.ones(dim), 0.3 * np.ones(nc - dim)]
    vals, grads = cb.mmd_se_batch(a, b, 0.4, coords, backward=backward, ctx=ctx)
    ovals, ograds = oracle.mmd_se_batch(a, b, 0.4, coords, backward=backward)
    tol = 2e-5 if ctx.precision == "fp32" else 1e-11
    assert np.allclose(vals, ovals, rtol=0, atol=tol * 2.0)
    assert np.allclose(grads, ograds, rtol=0, atol=tol * np.abs(ograds).max() * 20 + 1e-15)


# ------------------------------------------------------------------------------------------------ KDE evaluate / bandwidth / sample
@pytest.mark.parametrize("d,N,Q", [(1, 10, 5), (2, 100, 333), (3, 5000, 1100), (6, 777, 64)])
def test_kde_eval_matches_oracle(cb, ctx, d, N, Q):
    rng = np.random.default_rng(N)
    mu = rng.normal(size=(N, d)) * 3 + 50.0
    sig = rng.uniform(0.2, 1.5, size=d)
    q = rng.normal(size=(Q, d)) * 3 + 50.0
    w = rng.uniform(0.1, 1, size=N)
    topo = [0] * d
    for weights in (None, w):
        m = cb.MKD(topo, mu, sig, weights)
        want = oracle.kde_eval(mu, sig, q, weights)
        got = m(q, ctx=ctx)
        assert np.allclose(got, want, rtol=RTOL[ctx.precision] * 3, atol=1e-300)
        gotl = m(q, logp=True, ctx=ctx)
        assert np.allclose(gotl, oracle.kde_eval(mu, sig, q, weights, logp=True), rtol=0, atol=RTOL[ctx.precision] * 30)


def test_bandwidth_matches_oracle_and_upstream_fixtures(cb, ctx, tut3):
    names = sorted(tut3)[::4]
    got = cb.kde_bandwidth_batch([tut3[n][0] for n in names], "Point2", ctx=ctx)
    for n, g in zip(names, got):
        want = oracle.kde_bw_loo(tut3[n][0])
        # FP64 follows the oracle's golden-section path exactly; FP32 may leave it near convergence (tol 1e-2)
        assert np.allclose(g, want, rtol=1e-6 if ctx.precision == "fp64" else 2e-2), (n, g, want)
        assert np.abs(g / tut3[n][1] - 1).max() < (1e-2 if ctx.precision == "fp64" else 3e-2)  # upstream's own output
    rng = np.random.default_rng(1)
    x = np.c_[rng.normal(size=3000), np.angle(np.exp(1j * (3.0 + 0.4 * rng.normal(size=3000))))]
    got = cb.kde_bandwidth(x, (0, 1), ctx=ctx)
    want = oracle.kde_bw_loo(x, circular=[0, 1])
    assert np.allclose(got, want, rtol=1e-6 if ctx.precision == "fp64" else 2e-2)
    m = cb.manikde("Point2", tut3["l2_3"][0], ctx=ctx)
    assert np.allclose(m.bw, got if False else cb.kde_bandwidth(tut3["l2_3"][0], ctx=ctx))


def test_sample_matches_oracle_stream(cb, ctx):
    rng = np.random.default_rng(2)
    mu, sig = rng.normal(size=(200, 3)) * [5, 5, 1], np.array([0.3, 0.2, 0.4])
    ub = 53 if ctx.precision == "fp64" else 24
    for w in (None, rng.uniform(0.1, 1, size=200)):
        m = cb.MKD("Pose2", mu, sig, w)
        got, = cb.sample(m, 5000, seed=77, stream=3, ctx=ctx)
        want = oracle.kde_sample(mu, sig, 5000, seed=77, stream=3, w=w, circular=[0, 0, 1], ubits=ub)
        assert np.allclose(got, want, atol=1e-9)
    assert np.abs(got[:, 2]).max() <= np.pi + 1e-12


# ------------------------------------------------------------------------------------------------ products
@pytest.mark.parametrize("N", [1, 2, 3, 37, 64, 100, 1000])
def test_device_tree_equals_oracle_tree(cb, ctx, N):
    rng = np.random.default_rng(N)
    pts = rng.normal(size=(N, 3)) * [4, 1, 0.3] + [100, 0, 0]
    bw = np.array([0.2, 0.1, 0.05])
    w = rng.uniform(0.5, 1.5, size=N) if N % 2 else None
    depth, perm, levels = cb.tree_levels(pts, bw, w, ctx=ctx)
    t = oracle.Tree(pts, bw, w)
    assert depth == t.depth and np.array_equal(perm, t.perm)
    for l in range(depth + 1):
        assert levels[l][0].shape[0] == t.cnt[l]
        assert np.allclose(levels[l][0], t.mean[l], rtol=1e-12, atol=1e-12)
        assert np.allclose(levels[l][1], t.var[l], rtol=1e-9, atol=1e-14)
        assert np.allclose(levels[l][2], t.wt[l], rtol=1e-12)


def _messages(rng, D, N, d, spread=1.0):
    dens = []
    for j in range(D):
        c = rng.normal(size=d) * 0.5 * spread
        pts = c + rng.normal(size=(N + 7 * j, d)) * rng.uniform(0.5, 1.5, size=d) * spread
        dens.append((pts, 1.06 * pts.std(0) * len(pts) ** -0.2))
    return dens


@pytest.mark.parametrize("D,N,d,circ", [(2, 100, 2, (0, 0)), (3, 100, 3, (0, 0, 1)), (4, 257, 6, (0, 0, 0, 1, 1, 1)),
                                        (2, 64, 1, (0,)), (8, 128, 3, (0, 0, 0)), (3, 50, 4, (0, 1, 0, 1))])
def test_product_fp64_reproduces_oracle_chains(cb, ctx64, D, N, d, circ):
    """Same tree, same Philox counters, FP64 arithmetic -> the device Gibbs chains must select the same labels as the
    oracle's (a chunked vs sequential cumulative sum can flip a draw only at the 1e-15 level)."""
    rng = np.random.default_rng(D * 100 + N)
    dens = _messages(rng, D, N, d, spread=0.6 if any(circ) else 1.0)
    Nout = 300
    pts, lab = cb.product_points(dens, circ, Nout, seed=5, stream=2, ctx=ctx64)
    opts, olab = oracle.product(dens, Nout, seed=5, circular=list(circ), stream=2, ubits=53)
    same = (lab == olab).all(axis=1)
    assert same.mean() >= 0.99, same.mean()
    assert np.allclose(pts[same], opts[same], rtol=1e-9, atol=1e-9)
    for j in range(D):
        assert lab[:, j].min() >= 0 and lab[:, j].max() < len(dens[j][0])


def two_sample_mmd(x, y, bw):
    return oracle.mmd(x, y, bw)


@pytest.mark.parametrize("D,N,d,circ", [(3, 100, 3, (0, 0, 1)), (4, 300, 6, (0, 0, 0, 1, 1, 1)), (2, 1000, 2, (0, 0))])
def test_product_fp32_matches_oracle_in_distribution(cb, ctx32, D, N, d, circ):
    rng = np.random.default_rng(D + N)
    dens = _messages(rng, D, N, d, spread=0.6)
    Nout = 2000
    pts, lab = cb.product_points(dens, circ, Nout, seed=11, ctx=ctx32)
    opts, olab = oracle.product(dens, Nout, seed=11, circular=list(circ), ubits=24)
    # most chains coincide even in FP32 (same uniforms), the rest are equally valid draws
    assert (lab == olab).all(axis=1).mean() > 0.9
    sd = opts.std(0)
    assert np.allclose(pts.mean(0), opts.mean(0), atol=0.1 * sd.max())
    assert np.allclose(np.cov(pts.T).reshape(d, d), np.cov(opts.T).reshape(d, d), atol=0.15 * (sd.max() ** 2))
    # fixed-seed two-sample test: MMD^2 between device and oracle sets vs the spread between two oracle sets
    opts2, _ = oracle.product(dens, Nout, seed=12, circular=list(circ), ubits=24)
    bwk = 0.5 / (sd**2).sum()
    assert two_sample_mmd(pts, opts, bwk) <= two_sample_mmd(opts2, opts, bwk) + 1e-3


def test_product_batch_is_independent_of_batching_and_sharding(cb, ctx):
    rng = np.random.default_rng(9)
    groups = [[cb.MKD("Pose2", *x) for x in _messages(rng, D, 90 + 10 * D, 3, 0.6)] for D in (2, 3, 4, 2)]
    batch = cb.manifoldProducts(groups, N=128, seed=3, streams=[4, 5, 6, 7], return_labels=True, bw=None, ctx=ctx)
    for b, g in enumerate(groups):
        solo = cb.manifoldProducts([g], N=128, seed=3, streams=[4 + b], return_labels=True, bw=None, ctx=ctx)[0]
        assert np.array_equal(batch[b][1], solo[1]) and np.array_equal(batch[b][0].pts, solo[0].pts)
        # output-sample sharding: samples [64, 128) computed alone equal the tail of the full product
        tail = cb.manifoldProducts([g], N=64, seed=3, streams=[4 + b], sample_offsets=[64], return_labels=True, bw=None, ctx=ctx)[0]
        assert np.array_equal(tail[1], solo[1][64:]) and np.array_equal(tail[0].pts, solo[0].pts[64:])


def test_product_batch_from_pinned_host_arrays_uses_the_single_launch_copies(cb, ctx, monkeypatch):
    """Batches of >= 64 arrays in pinned (mapped) host memory are gathered / scattered by kernels that read and write the caller's
    arrays in place (chunked on a second stream so that the upload hides behind the tree build); results must equal the per-array copy path bit for bit.  Nout * D is odd for the labels
    (4-byte tail of the 8-byte-word copy) and the bandwidth outputs are pageable (mixed batch)."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(12)
    B, D, N, Nout, d = 24, 3, 4000, 33, 3  # 6.9 MB of inputs: above the packed-staging limit of small batches
    L = cb._capi.lib()
    msgs = [[_messages(rng, 1, N, d, 0.5)[0] for _ in range(D)] for _ in range(B)]
    pin = [[torch.from_numpy(np.ascontiguousarray(m[0])).pin_memory() for m in g] for g in msgs]
    circ = np.array([0, 0, 1], dtype=np.uint8)

    def run():
        out = [torch.zeros((Nout, d), dtype=torch.float64).pin_memory() for _ in range(B)]
        lab = [torch.zeros((Nout, D), dtype=torch.int32).pin_memory() for _ in range(B)]
        bw = [np.zeros(d) for _ in range(B)]
        keep, descs = [], (cb.ProductDesc * B)()
        for b in range(B):
            arr = (cb.Density * D)(*[cb.Density(N, pin[b][j].data_ptr(), msgs[b][j][1].ctypes.data, None) for j in range(D)])
            keep.append(arr)
            descs[b].dens, descs[b].D, descs[b].Nout, descs[b].niter, descs[b].stream = arr, D, Nout, 1, b
            descs[b].pts_out, descs[b].labels_out, descs[b].bw_out = out[b].data_ptr(), lab[b].data_ptr(), bw[b].ctypes.data
        ctx.check(L.cb2_product_batch(ctx._h, descs, B, circ.ctypes.data_as(C.c_void_p), d, 77))
        return [o.numpy().copy() for o in out], [l.numpy().copy() for l in lab], [x.copy() for x in bw]

    launches0 = L.cb2_kernel_launches(ctx._h)
    direct = run()
    n_direct = L.cb2_kernel_launches(ctx._h) - launches0
    monkeypatch.setenv("CB2_NO_DIRECT", "1")
    launches0 = L.cb2_kernel_launches(ctx._h)
    plain = run()
    n_plain = L.cb2_kernel_launches(ctx._h) - launches0
    # gather and scatter kernels instead of per-array copies: four upload chunks on the copy stream (the trees of a chunk are built
    # as soon as it has landed: three more tree launches) and one scatter of the sample points and labels
    assert n_direct == n_plain + 8, (n_direct, n_plain)
    monkeypatch.setenv("CB2_NO_COPY_OVERLAP", "1")
    monkeypatch.delenv("CB2_NO_DIRECT")
    launches0 = L.cb2_kernel_launches(ctx._h)
    serial = run()
    assert L.cb2_kernel_launches(ctx._h) - launches0 == n_plain + 2  # one gather and one scatter launch on the main stream
    for a, b in zip(serial, plain):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for a, b in zip(direct, plain):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert all((l >= 0).all() and (l < N + 14).all() for l in direct[1]) and all((x > 0).all() for x in direct[2])


def test_manifold_product_api_behaviour(cb, ctx, tut3):
    a, b, c = (cb.MKD("Point2", *tut3[k]) for k in ("l3_2", "l3_3", "l3_4"))
    p = cb.manifoldProduct([a, b, c], seed=1, ctx=ctx)
    assert p.N == 100 and p.d == 2 and np.isfinite(p.pts).all() and (p.bw > 0).all()
    want_bw = oracle.kde_bw_loo(p.pts)
    assert np.allclose(p.bw, want_bw, rtol=1e-6 if ctx.precision == "fp64" else 2e-2)  # fresh LOO bandwidth on the result
    assert cb.manifoldProduct([a], ctx=ctx) is a                                        # single input is returned unchanged
    p2 = cb.manifoldProduct([a, b], N=250, seed=1, ctx=ctx)
    assert p2.N == 250
    # the ZMQ multiplyDistributions example (ext/services/additional.jl:34): two 10-point 1-D densities
    x1 = np.array([-0.8272759752535751, 1.90340735325013, 0.5617238720356695, -1.2610130840544629, -0.5196580171571705,
                   0.47808234574398156, -0.5034354386657888, -0.1289061606999511, -0.6283440136328498, -0.04520468574277743])
    x2 = np.array([0.6986754683541115, 0.4206784705438769, -0.01798641761511103, 0.2862258908507545, 0.5021835960347083,
                   -1.4424531814460273, 0.7667313775874391, 0.13048774810077268, -0.004778323705883176, -0.6408606329686699])
    f1, f2 = cb.manikde("ContinuousScalar", x1[:, None], ctx=ctx), cb.manikde("ContinuousScalar", x2[:, None], ctx=ctx)
    pq = cb.manifoldProduct([f1, f2], N=2000, seed=4, ctx=ctx)
    o, _ = oracle.product([(f1.pts, f1.bw), (f2.pts, f2.bw)], 2000, seed=4, ubits=53 if ctx.precision == "fp64" else 24)
    assert abs(pq.pts.mean() - o.mean()) < 0.05 and abs(pq.pts.std() - o.std()) < 0.05


def test_product_far_apart_messages_do_not_underflow(cb, ctx):
    """Upstream has no max-subtraction and can underflow; the device path must stay finite."""
    rng = np.random.default_rng(4)
    dens = [(rng.normal(size=(80, 2)) * 0.1, np.array([0.02, 0.02])), (rng.normal(size=(80, 2)) * 0.1 + 40.0, np.array([0.02, 0.02]))]
    pts, lab = cb.product_points(dens, (0, 0), 200, seed=1, ctx=ctx)
    assert np.isfinite(pts).all() and np.abs(pts.mean(0) - 20.0).max() < 1.0


def test_full_size_product_chains_match_oracle_subset(cb, ctx64):
    """C5 shape (D=8 messages x N=4096 kernels, d=6, 3 circular): the oracle re-runs a slice of the chains."""
    rng = np.random.default_rng(1000)
    circ = (0, 0, 0, 1, 1, 1)
    truth = np.r_[rng.normal(size=3) * 2, rng.uniform(-1, 1, size=3)]
    dens = []
    for j in range(8):
        off = np.r_[rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.1]
        comp = rng.uniform(size=4096) < 0.5
        pts = truth + off + np.where(comp[:, None], 0.4, -0.4) * np.r_[1, 0, 0, 0, 0, 0] + rng.normal(size=(4096, 6)) * np.r_[0.5, 0.5, 0.5, 0.15, 0.15, 0.15]
        dens.append((pts, 1.06 * pts.std(0) * 4096 ** -0.2))
    pts, lab = cb.product_points(dens, circ, 4096, seed=21, stream=9, ctx=ctx64)
    assert np.isfinite(pts).all() and np.abs(pts[:, 3:]).max() <= np.pi + 1e-12
    assert lab.min() >= 0 and lab.max() < 4096
    opts, olab = oracle.product(dens, 96, seed=21, circular=list(circ), stream=9, ubits=53, sample_offset=2000)
    same = (lab[2000:2096] == olab).all(axis=1)
    assert same.mean() >= 0.97
    assert np.allclose(pts[2000:2096][same], opts[same], atol=1e-8)


# ------------------------------------------------------------------------------------------------ ScatterAlign
def _blobs(rng, n_per=50):
    c = np.array([[3.0, 0], [8.0, 0], [0, 5.0]])
    return np.concatenate([ci + 0.1 * rng.normal(size=(n_per, 2)) for ci in c])


def test_scatter_align_pose2_recovers_transform(cb, ctx):
    """test/testScatterAlignPose2.jl:171-199: 3-blob clouds, true qTp offset [2, 0, pi/6]; the reference asserts
    recovery within 1.5 m / 0.2 rad (:68-69)."""
    rng = np.random.default_rng(0)
    p1 = _blobs(rng)
    p2 = _blobs(rng)
    true = np.array([2.0, 0.0, np.pi / 6])
    # cloud2 = independent redraw of the scene mapped by qTp (testScatterAlignPose2.jl:180-186): q = T(true) p
    q2 = oracle.transform_points(2, true, p2, backward=False)
    c1, c2 = cb.manikde("Point2", p1, ctx=ctx), cb.manikde("Point2", q2, ctx=ctx)
    sap = cb.ScatterAlignPose2(cloud1=c1, cloud2=c2, sample_count=100, bw=1.0)
    cf = cb.CalcFactor(sap, cb.preambleCache(None, [], sap))
    x0 = np.tile(true, (16, 1)) + np.c_[0.5 * rng.normal(size=(16, 2)), 0.1 * rng.normal(size=16)]
    coords, score, iters = cb.getSample(cf, x0=x0, seed=5, ctx=ctx, return_info=True)
    # cost = mmd(smps1, T(c)^-1 smps2) is minimal at c = true
    assert np.abs(coords[:, :2] - true[:2]).max() < 1.5 and np.abs(coords[:, 2] - true[2]).max() < 0.2
    assert (score < 0.5).all() and (iters > 0).all() and cf.cache["score"][0] == score[-1]
    # the optimiser's final cost is the device mmd at the optimum of its own sample sets: never above the start cost
    start = cb.getSample(cf, x0=x0, seed=5, max_iter=1, ctx=ctx, return_info=True)[1]
    assert (score <= start + 1e-12).all()


# ------------------------------------------------------------------------------------------------ error behaviour
def test_errors_are_codes_not_crashes(cb, ctx):
    with pytest.raises(cb.CB2Error) as e:
        cb.mmd(np.zeros((3, 9)), np.zeros((3, 9)), 1.0, ctx=ctx)
    assert e.value.code == -5
    with pytest.raises(cb.CB2Error) as e:
        cb.mmd(np.zeros((3, 2)), np.zeros((3, 2)), -1.0, ctx=ctx)
    assert e.value.code == -1 and "bw" in str(e.value)
    with pytest.raises(cb.CB2Error) as e:
        cb.product_points([(np.zeros((20000, 2)), np.ones(2))] * 2, (0, 0), 10, ctx=ctx)
    assert e.value.code == -5
    with pytest.raises(cb.CB2Error) as e:
        cb.product_points([(np.zeros((10, 2)), np.zeros(2))] * 2, (0, 0), 10, ctx=ctx)
    assert e.value.code == -1 and "bandwidth" in str(e.value)
    # the context is still usable afterwards
    assert np.isfinite(cb.mmd(np.zeros((3, 2)), np.ones((3, 2)), 1.0, ctx=ctx))


# ------------------------------------------------------------------------------------------------ limits
def test_product_at_compiled_limits(cb, ctx64):
    """Maximum sizes the library accepts: N = 16384 kernels per message, D = 16 messages, generic d = 5 mask path."""
    rng = np.random.default_rng(77)
    # N = 16384 (largest tree: 213 KB of shared memory, 15 levels), D = 2, d = 2; the oracle re-runs a slice of the chains
    dens = [(rng.normal(size=(16384, 2)) * [2, 1] + [j, 0], np.array([0.08, 0.05])) for j in range(2)]
    pts, lab = cb.product_points(dens, (0, 0), 512, seed=3, ctx=ctx64)
    opts, olab = oracle.product(dens, 64, seed=3, circular=[0, 0], ubits=53, sample_offset=300)
    same = (lab[300:364] == olab).all(axis=1)
    assert same.mean() >= 0.95 and np.allclose(pts[300:364][same], opts[same], atol=1e-8)
    depth, perm, levels = cb.tree_levels(dens[0][0], dens[0][1], ctx=ctx64)
    assert depth == 14 and sorted(perm.tolist()) == list(range(16384))
    # D = 16 messages, ragged sizes, d = 5 with a mixed circular mask (run-time mask kernel)
    dens = [(rng.normal(size=(40 + 3 * j, 5)) * 0.5, np.full(5, 0.3)) for j in range(16)]
    circ = (0, 1, 0, 0, 1)
    pts, lab = cb.product_points(dens, circ, 200, seed=4, ctx=ctx64)
    opts, olab = oracle.product(dens, 200, seed=4, circular=list(circ), ubits=53)
    assert (lab == olab).all(axis=1).mean() >= 0.98
    # one output sample, one kernel per message
    pts, lab = cb.product_points([(np.array([[1.0, 2.0]]), np.array([0.5, 0.5])), (np.array([[2.0, 2.0]]), np.array([0.5, 0.5]))],
                                 (0, 0), 1, seed=1, ctx=ctx64)
    assert pts.shape == (1, 2) and (lab == 0).all() and abs(pts[0, 0] - 1.5) < 2.0
    # bandwidth search at the largest supported N
    x = rng.normal(size=(16384, 1))
    assert np.allclose(cb.kde_bandwidth(x, ctx=ctx64), oracle.kde_bw_loo(x), rtol=1e-6)


def test_product_with_weighted_kernels(cb, ctx64, ctx32):
    """Per-kernel weights (the non-uniform leaf path): FP64 chains reproduce the oracle, FP32 agrees in distribution."""
    rng = np.random.default_rng(55)
    dens = []
    for j in range(3):
        pts = rng.normal(size=(150 + j, 3)) * [1.0, 0.7, 0.2] + [0.2 * j, 0, 0]
        w = rng.uniform(0.05, 1.0, size=len(pts))
        w[::7] = 0.0  # some kernels carry no mass at all
        dens.append((pts, np.array([0.3, 0.25, 0.08]), w))
    opts, olab = oracle.product(dens, 400, seed=8, circular=[0, 0, 1], ubits=53)
    pts, lab = cb.product_points(dens, (0, 0, 1), 400, seed=8, ctx=ctx64)
    same = (lab == olab).all(axis=1)
    assert same.mean() >= 0.99 and np.allclose(pts[same], opts[same], atol=1e-9)
    for j in range(3):
        assert (dens[j][2][lab[:, j]] > 0).all()  # zero-weight kernels are never selected
    pts32, lab32 = cb.product_points(dens, (0, 0, 1), 400, seed=8, ctx=ctx32)
    o24, l24 = oracle.product(dens, 400, seed=8, circular=[0, 0, 1], ubits=24)
    assert (lab32 == l24).all(axis=1).mean() > 0.9 and np.allclose(pts32.mean(0), o24.mean(0), atol=0.1)


def test_product_random_shapes_reproduce_oracle(cb, ctx64):
    """Randomised shape sweep (ragged message sizes, every dimension 1..6, random circular masks, optional weights,
    Niter 0..2): FP64 device chains must select the oracle's labels."""
    rng = np.random.default_rng(2024)
    for case in range(24):
        d = int(rng.integers(1, 7))
        D = int(rng.integers(2, 7))
        circ = tuple(int(c) for c in rng.integers(0, 2, size=d))
        dens = []
        for j in range(D):
            N = int(rng.choice([1, 2, 3, 5, 17, 64, 100, 129, 300]))
            pts = rng.normal(size=(N, d)) * rng.uniform(0.2, 1.0, size=d) + rng.normal(size=d) * 0.3
            for i, c in enumerate(circ):
                if c:
                    pts[:, i] = (pts[:, i] + rng.uniform(-3, 3) + np.pi) % (2 * np.pi) - np.pi  # may straddle the cut
            bw = rng.uniform(0.1, 0.5, size=d)
            w = rng.uniform(0.1, 1.0, size=N) if rng.uniform() < 0.3 else None
            dens.append((pts, bw, w))
        Nout, niter = int(rng.choice([1, 7, 130, 257])), int(rng.integers(0, 3))
        pts, lab = cb.product_points(dens, circ, Nout, seed=case, niter=niter, stream=case, ctx=ctx64)
        opts, olab = oracle.product(dens, Nout, seed=case, circular=list(circ), niter=niter, stream=case, ubits=53)
        same = (lab == olab).all(axis=1)
        assert same.mean() >= 0.97, (case, d, D, circ, [len(x[0]) for x in dens], Nout, niter, same.mean())
        assert np.allclose(pts[same], opts[same], atol=1e-8), case


def test_pair_kernels_random_shapes(cb, ctx):
    """Randomised sizes around the tile boundaries (1024 rows, 512/4096 columns) for mmd, evaluate, bandwidth."""
    rng = np.random.default_rng(99)
    tol = RTOL[ctx.precision]
    for case in range(10):
        d = int(rng.integers(1, 7))
        N, M = (int(rng.choice([1, 2, 511, 512, 513, 1023, 1024, 1025, 4095, 4097, 6000])) for _ in range(2))
        a, b = rng.normal(size=(N, d)) + 3.0, rng.normal(size=(M, d)) * 0.8 + 3.2
        bw = float(rng.uniform(0.05, 2.0))
        _, sums = oracle.mmd(a, b, bw, return_sums=True)
        assert np.allclose(cb.mmd_sums(a, b, bw, ctx=ctx), sums, rtol=tol), (case, d, N, M)
        sig = rng.uniform(0.3, 1.2, size=d)
        Q = int(rng.choice([1, 300, 1025]))
        q = rng.normal(size=(Q, d)) + 3.0
        w = rng.uniform(0.2, 1.0, size=N) if case % 2 else None
        got, want = cb.MKD([0] * d, a, sig, w)(q, ctx=ctx), oracle.kde_eval(a, sig, q, w)
        bulk = want > 1e-12 * want.max()
        assert np.allclose(got[bulk], want[bulk], rtol=3 * tol), (case, d, N, Q)
        # far tails: the FP32 exponent itself carries ~|log2 p| * 2^-24 of error, so compare in the log domain there
        tail = ~bulk & (want > (1e-30 if ctx.precision == "fp32" else 1e-300))  # FP32 sums flush to zero below 2^-126
        assert np.allclose(np.log(np.maximum(got[tail], 1e-300)), np.log(want[tail]), rtol=0, atol=1e-3 if ctx.precision == "fp32" else 1e-9)
    for N in (2, 3, 100, 1024, 1025, 2500):
        x = np.c_[rng.normal(size=N) * 2.0, rng.uniform(-np.pi, np.pi, size=N)]
        got, want = cb.kde_bandwidth(x, (0, 1), ctx=ctx), oracle.kde_bw_loo(x, circular=[0, 1])
        assert np.allclose(got, want, rtol=1e-6 if ctx.precision == "fp64" else 3e-2), (N, got, want)


def test_device_resident_belief_round_trip_and_packed_format(cb, ctx, tut3):
    """a tut3 belief uploaded to the device packs back into exactly the PackedManifoldKernelDensity JSON of the
    host object (the fixture format), and a product of resident beliefs equals the host-pointer product bit for bit."""
    import json
    from importlib import import_module
    res = import_module("caesar_jl_b200.resident")
    slab = res.DeviceSlab(ctx, 8 << 20)
    host = [cb.MKD("Point2", *tut3[k]) for k in ("l3_2", "l3_3", "l3_4")]
    dev = [res.upload_belief(m, slab, ctx) for m in host]
    for m, dm in zip(host, dev):
        assert json.dumps(cb.packDistribution(dm), sort_keys=True) == json.dumps(cb.packDistribution(m), sort_keys=True)
    want = cb.manifoldProducts([host, host[:2]], "Point2", N=100, seed=11, streams=[3, 4], ctx=ctx)
    got = res.manifoldProducts_dev([dev, dev[:2]], "Point2", 100, slab, seed=11, streams=[3, 4], ctx=ctx)
    for w, g in zip(want, got):
        h = g.to_host()
        assert np.array_equal(h.pts, w.pts) and np.array_equal(h.bw, w.bw)
    # fresh bandwidths of resident sample sets = the host call's
    arrs = [d.pts for d in dev]
    bel = res.manikde_dev("Point2", arrs, slab, ctx)
    ref = cb.kde_bandwidth_batch([m.pts for m in host], "Point2", ctx=ctx)
    for b, r in zip(bel, ref):
        assert np.array_equal(b.to_host().bw, r)
    slab.free()


@pytest.mark.parametrize("circ", [(0, 0, 0, 0, 0, 0), (0, 0, 0, 1, 1, 1), (1, 0, 1, 0, 1, 0)])
@pytest.mark.parametrize("weighted", [False, True])
def test_product_fp32_six_dimensional_packed_paths(cb, ctx32, circ, weighted):
    """The FP32 d = 6 kernels score with packed f32x2 arithmetic whenever a warp needs no wrap (run-time mask kernel for the
    non-Pose3 masks, uniform and per-kernel weights, ragged sizes, beliefs that do / do not straddle the +-pi cut): the chains
    must still follow the oracle's (same uniforms; FP32 rounding flips a few draws)."""
    rng = np.random.default_rng(77 + sum(circ) + 10 * weighted)
    for straddle in (False, True):
        dens = []
        for j, N in enumerate((257, 300, 64, 129)):
            pts = rng.normal(size=(N, 6)) * rng.uniform(0.2, 0.6, size=6) + rng.normal(size=6) * 0.2
            for i, c in enumerate(circ):
                if c:
                    pts[:, i] = (pts[:, i] + (3.0 if straddle else 0.5) + np.pi) % (2 * np.pi) - np.pi
            w = rng.uniform(0.1, 1.0, size=N) if weighted else None
            dens.append((pts, rng.uniform(0.15, 0.4, size=6), w))
        pts, lab = cb.product_points(dens, circ, 600, seed=21, niter=1, ctx=ctx32)
        opts, olab = oracle.product(dens, 600, seed=21, circular=list(circ), niter=1, ubits=24)
        assert np.isfinite(pts).all()
        assert (lab == olab).all(axis=1).mean() > 0.9, (circ, weighted, straddle, (lab == olab).all(axis=1).mean())
        same = (lab == olab).all(axis=1)
        diff = pts[same] - opts[same]
        for i, c in enumerate(circ):
            if c:
                diff[:, i] = (diff[:, i] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(diff).max() < 5e-5   # same labels, same normal draws: FP32 arithmetic on the final mean only


@pytest.mark.parametrize("D,N", [(2, 64), (2, 100), (3, 1024), (2, 4096), (8, 513)])
def test_product_fp32_tensor_core_leaf_pass_follows_the_oracle_chains(cb, ctx32, D, N):
    """FP32 d = 6 wrap-free leaf passes run on the tensor cores (tcgen05.mma kind::tf32 with a 3 x TF32 split, accumulators in
    TMEM; gibbs.cuh leaf_pass_tc): tile counts 1, partial, many; the chains must still select the oracle's labels (same uniforms,
    scores equal to ~1e-6) -- including Pose3 beliefs away from the +-pi cut and per-kernel weights."""
    rng = np.random.default_rng(100 + D + N)
    for circ, weighted in (((0,) * 6, False), ((0, 0, 0, 1, 1, 1), True)):
        dens = []
        for j in range(D):
            pts = rng.normal(size=(N, 6)) * 0.5 + rng.normal(size=6) * 0.2
            w = rng.uniform(0.2, 1.0, size=N) if weighted else None
            dens.append((pts, np.full(6, 0.3) if N < 1000 else np.full(6, 0.2), w))
        pts, lab = cb.product_points(dens, circ, 384, seed=3, ctx=ctx32)
        opts, olab = oracle.product(dens, 384, seed=3, circular=list(circ), ubits=24)
        same = (lab == olab).all(axis=1)
        assert same.mean() > 0.97, (D, N, circ, weighted, same.mean())
        assert np.abs(pts[same] - opts[same]).max() < 5e-5


# ------------------------------------------------------------------------------------------------ round-2 additions
@pytest.mark.parametrize("d,circ", [(4, (0, 1, 0, 1)), (5, (0, 0, 0, 1, 1)), (4, (0, 0, 0, 0)), (5, (1, 0, 0, 0, 0))])
@pytest.mark.parametrize("N", [512, 4096])
def test_product_fp64_d4_d5_many_tiles_reproduce_oracle_chains(cb, ctx64, d, circ, N):
    """d = 4, 5 have 48-byte (FP32) / 96-byte (FP64) coarse records, so a tile holds 170 of them: not a whole number of scoring
    groups.  Levels with more than one tile (N >= 257) must still close their chunk sums on the group grid (gibbs.cuh TN_NODE)."""
    rng = np.random.default_rng(40 + d + N)
    dens = _messages(rng, 3, N, d, spread=0.5 if any(circ) else 1.0)
    Nout = 320
    pts, lab = cb.product_points(dens, circ, Nout, seed=17, stream=1, ctx=ctx64)
    opts, olab = oracle.product(dens, Nout, seed=17, circular=list(circ), stream=1, ubits=53)
    same = (lab == olab).all(axis=1)
    assert same.mean() >= 0.99, same.mean()
    assert np.allclose(pts[same], opts[same], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("d", [4, 5])
def test_product_fp32_d4_d5_many_tiles_follow_oracle_chains(cb, ctx32, d):
    rng = np.random.default_rng(50 + d)
    circ = (0, 0, 0, 1, 1)[:d]
    dens = _messages(rng, 3, 1500, d, spread=0.5)
    pts, lab = cb.product_points(dens, circ, 512, seed=23, ctx=ctx32)
    opts, olab = oracle.product(dens, 512, seed=23, circular=list(circ), ubits=24)
    assert (lab == olab).all(axis=1).mean() > 0.9


def _c5_variable(seed):
    """One variable of the bench workload (bench.make_variable): D = 8 messages x N = 4096 kernels, Pose3."""
    import bench
    return bench.make_variable(seed)


def test_headline_instance_fp32_c5_shape_follows_oracle_chains(cb, ctx32):
    """The exact kernel instance bench.py times -- FP32, D = 8 x N = 4096 x N_out = 4096, d = 6 Pose3, packed coarse levels and the
    tensor-core leaf pass over 32 tiles -- against the oracle on 256 of its chains (same Philox uniforms, ubits = 24)."""
    circ = (0, 0, 0, 1, 1, 1)
    dens = _c5_variable(3)
    pts, lab = cb.product_points(dens, circ, 4096, seed=31, stream=3, ctx=ctx32)
    assert np.isfinite(pts).all() and lab.min() >= 0 and lab.max() < 4096
    for off in (0, 3840):  # first and last CTA of the product
        opts, olab = oracle.product(dens, 128, seed=31, circular=list(circ), stream=3, ubits=24, sample_offset=off)
        sl = slice(off, off + 128)
        same = (lab[sl] == olab).all(axis=1)
        assert same.mean() > 0.9, (off, same.mean())
        assert (lab[sl] == olab).mean() > 0.97   # per (chain, message) label
        diff = pts[sl][same] - opts[same]
        diff[:, 3:] = (diff[:, 3:] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(diff).max() < 5e-5
    # in distribution against the FP64 device chains of the same seeds
    ctx64 = cb.Context(0, "fp64")
    p64, _ = cb.product_points(dens, circ, 4096, seed=31, stream=3, ctx=ctx64)
    dm = p64.mean(0) - pts.mean(0)
    dm[3:] = (dm[3:] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(dm).max() < 0.02


def test_product_dimension_above_six_is_rejected_before_any_work(cb, ctx):
    rng = np.random.default_rng(0)
    dens = [(rng.normal(size=(64, 7)), np.full(7, 0.3)) for _ in range(2)]
    launches = ctx.kernel_launches
    with pytest.raises(cb.CB2Error) as ei:
        cb.product_points(dens, (0,) * 7, 16, seed=1, ctx=ctx)
    assert ei.value.code == -5 and ctx.kernel_launches == launches


def test_tensor_core_leaf_pass_second_pass_consistency_is_quantified(cb, ctx32):
    """With the tensor-core leaf pass the first pass scores leaves with a 3 x TF32 product and the second pass (re-sum of the
    selected chunk) with FP32 FMAs: equal to ~1e-6, not bit for bit.  A draw whose target sits within that of a chunk boundary ends
    its re-sum short and takes the chunk's last candidate with mass -- the neighbour of the exact draw.  Count how often."""
    dens = _c5_variable(5)
    circ = (0, 0, 0, 1, 1, 1)
    ctx32.debug_counter(0, reset=True)
    pts, lab = cb.product_points(dens, circ, 4096, seed=41, stream=1, ctx=ctx32)
    short = ctx32.debug_counter(0, reset=True)
    draws = 4096 * 8 * 2 * 13  # chains x messages x sweeps x level visits
    assert short / draws < 2e-4, (short, draws)
    # the chains still follow the oracle's
    opts, olab = oracle.product(dens, 128, seed=41, circular=list(circ), stream=1, ubits=24)
    assert (lab[:128] == olab).mean() > 0.97


@pytest.mark.parametrize("N", [192, 500, 1000, 1024])
def test_bandwidth_cluster_kernel_matches_single_cta_kernel_and_oracle(cb, ctx, N, monkeypatch):
    """Small batches of a few hundred points take the thread-block-cluster search (eight CTAs per coordinate, partial sums
    exchanged through distributed shared memory); it must walk the same golden-section path as the single-CTA kernel."""
    rng = np.random.default_rng(N)
    pts = np.c_[rng.normal(size=(N, 2)) * [3.0, 0.5] + (rng.uniform(size=(N, 1)) < 0.4) * [6.0, 0.0], rng.normal(size=N) * 0.4 + 3.0]
    pts[:, 2] = (pts[:, 2] + np.pi) % (2 * np.pi) - np.pi   # heading straddles the +-pi cut
    launches = ctx.kernel_launches
    got = cb.kde_bandwidth(pts, "Pose2", ctx=ctx)
    assert ctx.kernel_launches > launches
    monkeypatch.setenv("CB2_NO_LOO_CLUSTER", "1")
    single = cb.kde_bandwidth(pts, "Pose2", ctx=ctx)
    monkeypatch.delenv("CB2_NO_LOO_CLUSTER")
    want = oracle.kde_bw_loo(pts, circular=[0, 0, 1])
    if ctx.precision == "fp64":
        assert np.allclose(got, single, rtol=1e-9) and np.allclose(got, want, rtol=1e-6)
    else:
        assert np.allclose(got, single, rtol=2e-2) and np.allclose(got, want, rtol=3e-2)


@pytest.mark.parametrize("N", [2, 3, 31, 32, 33, 64, 65, 96, 97, 100, 128])
def test_bandwidth_tiny_kernel_matches_single_cta_kernel_and_oracle(cb, ctx, N, monkeypatch):
    """N <= 128 (the solver's default 100 particles) takes loo_tiny_kernel: (row, column-segment) thread layout, per-thread copies
    of the search state, bandwidths written straight to the output.  Same golden-section path as the general single-CTA kernel."""
    rng = np.random.default_rng(1000 + N)
    pts = np.c_[rng.normal(size=(N, 2)) * [3.0, 0.5] + (rng.uniform(size=(N, 1)) < 0.4) * [6.0, 0.0], rng.normal(size=N) * 0.4 + 3.0]
    pts[:, 2] = (pts[:, 2] + np.pi) % (2 * np.pi) - np.pi   # heading straddles the +-pi cut: the wrapped loop
    launches = ctx.kernel_launches
    got = cb.kde_bandwidth(pts, "Pose2", ctx=ctx)
    n_tiny = ctx.kernel_launches - launches
    monkeypatch.setenv("CB2_NO_LOO_TINY", "1")
    launches = ctx.kernel_launches
    single = cb.kde_bandwidth(pts, "Pose2", ctx=ctx)
    assert ctx.kernel_launches - launches == n_tiny + 1   # the general path also launches loo_collect_kernel
    monkeypatch.delenv("CB2_NO_LOO_TINY")
    want = oracle.kde_bw_loo(pts, circular=[0, 0, 1])
    if ctx.precision == "fp64":
        assert np.allclose(got, single, rtol=1e-9) and np.allclose(got, want, rtol=1e-6)
    else:
        assert np.allclose(got, single, rtol=2e-2) and np.allclose(got, want, rtol=3e-2)
    # a batch that mixes tiny and larger densities goes through both kernels and the collect launch
    big = rng.normal(size=(300, 3))
    both = cb.kde_bandwidth_batch([pts, big], "Pose2", ctx=ctx)
    assert np.allclose(both[0], got, rtol=1e-12) and np.allclose(both[1], cb.kde_bandwidth(big, "Pose2", ctx=ctx), rtol=1e-12)


@pytest.mark.parametrize("N,d", [(2, 2), (3, 3), (100, 3), (100, 2), (129, 6), (255, 3), (256, 6), (200, 5)])
def test_small_tree_build_equals_general_tree_build(cb, ctx, N, d, monkeypatch):
    """Messages of N <= 256 points build their tree in shared memory (rank ordering inside the nodes instead of the bitonic
    network): same permutation and statistics as the general kernel, bit for bit, and the same tree as the oracle."""
    rng = np.random.default_rng(7 * N + d)
    pts = rng.normal(size=(N, d)) * rng.uniform(0.3, 4.0, size=d)
    pts[: N // 3] = pts[N // 3: 2 * (N // 3)]      # duplicated points: ties are broken by position
    bw = rng.uniform(0.05, 0.3, size=d)
    w = rng.uniform(0.5, 1.5, size=N) if N % 2 else None
    depth, perm, levels = cb.tree_levels(pts, bw, w, ctx=ctx)
    monkeypatch.setenv("CB2_NO_TREE_SMALL", "1")
    depth2, perm2, levels2 = cb.tree_levels(pts, bw, w, ctx=ctx)
    monkeypatch.delenv("CB2_NO_TREE_SMALL")
    assert depth == depth2 and np.array_equal(perm, perm2)
    for a, b in zip(levels, levels2):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    if d <= 6:
        t = oracle.Tree(pts, bw, w)
        assert depth == t.depth and np.array_equal(perm, t.perm)
        assert np.allclose(levels[0][0], t.mean[0], rtol=1e-12, atol=1e-12)
    if d <= 6:   # and a product through it: identical chains either way
        dens = [(pts, bw), (pts[::-1] * 1.01 + 0.05, bw * 1.3)]
        a = cb.product_points(dens, (0,) * d, 300, seed=4, ctx=ctx)
        monkeypatch.setenv("CB2_NO_TREE_SMALL", "1")
        b = cb.product_points(dens, (0,) * d, 300, seed=4, ctx=ctx)
        monkeypatch.delenv("CB2_NO_TREE_SMALL")
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("manifold,D,N", [("Point2", 2, 2), ("Point2", 3, 100), ("Pose2", 3, 100), ("Pose2", 2, 255), ("Pose2", 5, 77),
                                           ("Pose2", 3, 300), ("Point2", 8, 128)])
def test_small_product_kernel_equals_general_kernel(cb, ctx, manifold, D, N, monkeypatch):
    """Planar products whose records fit in shared memory run the SMALL instance of the Gibbs kernel (all levels staged once, no
    bulk copies, second pass and leave-one-out gathers from shared memory): identical labels and points."""
    d = 2 if manifold == "Point2" else 3
    circ = (0, 0) if d == 2 else (0, 0, 1)
    rng = np.random.default_rng(11 * N + D)
    dens = []
    for j in range(D):
        pts = rng.normal(size=(N + 3 * j, d)) * rng.uniform(0.5, 1.5, size=d) + rng.normal(size=d) * 0.3
        if d == 3:
            pts[:, 2] = (pts[:, 2] * 0.3 + 3.0 + np.pi) % (2 * np.pi) - np.pi   # straddles the cut
        w = rng.uniform(0.2, 1.0, size=len(pts)) if j % 2 else None
        dens.append((pts, 1.06 * pts.std(0) * len(pts) ** -0.2) + ((w,) if w is not None else ()))
    a = cb.product_points(dens, circ, 333, seed=9, ctx=ctx)
    monkeypatch.setenv("CB2_NO_GIBBS_SMALL", "1")
    b = cb.product_points(dens, circ, 333, seed=9, ctx=ctx)
    monkeypatch.delenv("CB2_NO_GIBBS_SMALL")
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


@pytest.mark.parametrize("scale", [1e-3, 1.0, 3e4, "mixed"])
def test_product_fp32_coarse_levels_common_denominator_and_its_range_guard(cb, ctx32, scale):
    """Coarse levels of the FP32 Pose3 kernel score over ONE common denominator of all six coordinates when a warp can prove that the
    six-fold product of variances stays inside FP32 range, and over two denominators of three otherwise.  Millimetre / milliradian
    beliefs (product of variances ~1e-36), kilometre-scale ones (1e+50 without the guard) and mixed scales must all follow the
    oracle's chains and stay finite."""
    rng = np.random.default_rng(5)
    sc = np.array([1e-3, 1e4, 1.0, 1e-3, 0.3, 1e-2]) if scale == "mixed" else np.r_[np.full(3, float(scale)), np.full(3, min(float(scale), 0.3))]
    circ = (0, 0, 0, 1, 1, 1)
    dens = []
    for j, N in enumerate((512, 300, 257, 1024)):
        pts = rng.normal(size=(N, 6)) * sc * rng.uniform(0.5, 1.0, size=6) + rng.normal(size=6) * 0.3 * sc
        dens.append((pts, 1.06 * pts.std(0) * N ** -0.2))
    pts, lab = cb.product_points(dens, circ, 512, seed=13, ctx=ctx32)
    opts, olab = oracle.product(dens, 512, seed=13, circular=list(circ), ubits=24)
    assert np.isfinite(pts).all()
    agree = (lab == olab).all(axis=1)
    assert agree.mean() > 0.9, (scale, agree.mean())
    diff = pts[agree] - opts[agree]
    diff[:, 3:] = (diff[:, 3:] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(diff / sc).max() < 1e-3
