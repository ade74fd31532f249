"""caesar.jl_b200 -- host-side mirror (Python, because Julia is absent from this image) of the Julia API that
Caesar.jl's belief hot path sits behind: `manikde!`, `manifoldProduct`, `mmd`, `sample`, `getPoints`, `getBW`,
evaluation of a ManifoldKernelDensity, and the ScatterAlignPose2/3 factor plugin.  Every operator is a thin call
into libcaesar_b200.so through the C ABI of include/caesar_b200.h -- the same entry points the Julia shim
(caesar.jl_b200/julia/CaesarB200.jl) binds with `ccall`.  No arithmetic of the hot path happens in Python and
there is no CPU fallback.

Reference interfaces mirrored (R: = /root/reference/):
  manikde!(M, pts; bw)                 call sites R:ext/Images/ScatterAlignPose2.jl:105-106, R:test/testScatterAlignPose2.jl:187
  manifoldProduct(ff, M; Niter, N)     call sites R:ext/services/additional.jl:24, R:docs/src/principles/multiplyingDensities.md:25
  mmd(M, a, b, N, M, threads; bw)      call site  R:ext/Images/ScatterAlignPose2.jl:179
  mmd(mkdA, mkdB), isapprox(;mmd_tol)  call sites R:test/ICRA2022_tests/insufficient_data.jl:53, R:test/testScatterAlignPose2.jl:81-82
  (mkd)(pts)                           call sites R:test/ICRA2022_tests/multihypo_corners.jl:66-90
  sample(mkd, n)                       call sites R:ext/Images/ScatterAlignPose2.jl:140-141,172-173
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import CB2_FP32, CB2_FP64, CB2Error, Density, ProductDesc, ConvDesc, f64, ptr, circ_arr

# variable types -> per-coordinate topology: 1 = circular (wrap to (-pi, pi])
MANIFOLDS = {
    "ContinuousScalar": (0,), "Circular": (1,), "Point2": (0, 0), "Point3": (0, 0, 0), "Position3": (0, 0, 0),
    "Pose2": (0, 0, 1), "Pose3": (0, 0, 0, 1, 1, 1),
    "TranslationGroup(2)": (0, 0), "TranslationGroup(3)": (0, 0, 0),
    "SpecialEuclidean(2)": (0, 0, 1), "SpecialEuclidean(3)": (0, 0, 0, 1, 1, 1),
}


def _topology(M, d=None):
    if M is None:
        return tuple([0] * d)
    if isinstance(M, str):
        key = M.replace("RoME.", "").replace("IncrementalInference.", "")
        if key not in MANIFOLDS:
            raise KeyError(f"unknown variable type / manifold {M!r}")
        return MANIFOLDS[key]
    return tuple(int(bool(c)) for c in M)


class Context:
    """One library context = one GPU (one process per GPU).  precision: 'fp32' (default) or 'fp64'."""

    def __init__(self, device=0, precision="fp32"):
        self._h = C.c_void_p()
        prec = {"fp32": CB2_FP32, "fp64": CB2_FP64, CB2_FP32: CB2_FP32, CB2_FP64: CB2_FP64}[precision]
        L = _capi.lib()
        rc = L.cb2_create(int(device), prec, C.byref(self._h))
        if rc != 0:
            raise CB2Error(rc, L.cb2_last_error(None).decode())
        self.device = int(device)
        self.precision = "fp64" if prec == CB2_FP64 else "fp32"

    def check(self, rc):
        if rc != 0:
            raise CB2Error(rc, _capi.lib().cb2_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _capi.lib().cb2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self.check(_capi.lib().cb2_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synchronize(self):
        self.check(_capi.lib().cb2_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(_capi.lib().cb2_kernel_launches(self._h))

    @property
    def sm_count(self):
        return int(_capi.lib().cb2_device_sm_count(self._h))

    def set_level_down(self, mode):
        """1 (default): weight-proportional random child on the way down the trees; 0: always the first child."""
        self.check(_capi.lib().cb2_set_level_down(self._h, int(mode)))

    def debug_counter(self, which=0, reset=False):
        v = C.c_uint64()
        self.check(_capi.lib().cb2_debug_counter(self._h, int(which), int(bool(reset)), C.byref(v)))
        return int(v.value)

    def stage_ms(self, name):
        return float(_capi.lib().cb2_last_stage_ms(self._h, name.encode()))


_default_ctx = {}


def context(device=0, precision="fp32"):
    key = (int(device), precision)
    if key not in _default_ctx:
        _default_ctx[key] = Context(device, precision)
    return _default_ctx[key]


class ManifoldKernelDensity:
    """Equally (or explicitly) weighted Gaussian-kernel density on a product of lines and circles.
    pts: N x d coordinates (row = point; memory-identical to Julia's d x N column-major matrix)."""

    def __init__(self, manifold, pts, bw, weights=None, partial=None, infoPerCoord=None):
        self.pts = f64(np.atleast_2d(pts))
        self.N, self.d = self.pts.shape
        self.manifold = manifold
        self.circular = _topology(manifold, self.d)
        if len(self.circular) != self.d:
            raise ValueError(f"manifold {manifold!r} has {len(self.circular)} coordinates, points have {self.d}")
        self.bw = f64(np.broadcast_to(np.asarray(bw, dtype=np.float64), (self.d,)))
        self.weights = None if weights is None else f64(weights)
        self.partial = partial
        self.infoPerCoord = np.ones(self.d) if infoPerCoord is None else f64(infoPerCoord)

    def __call__(self, q, logp=False, ctx=None):
        return evaluate(self, q, logp=logp, ctx=ctx)

    def _density(self):
        return Density(self.N, self.pts.ctypes.data, self.bw.ctypes.data,
                       self.weights.ctypes.data if self.weights is not None else None)


MKD = ManifoldKernelDensity


def getPoints(mkd):
    return mkd.pts


def getBW(mkd):
    return mkd.bw


def Npts(mkd):
    return mkd.N


def Ndim(mkd):
    return mkd.d


def kde_bandwidth(pts, manifold=None, ctx=None):
    """Leave-one-out likelihood cross-validation bandwidth, one sigma per coordinate (what `manikde!` computes
    when `bw` is not given)."""
    ctx = ctx or context()
    pts = f64(np.atleast_2d(pts))
    N, d = pts.shape
    circ = circ_arr(_topology(manifold, d), d)
    out = np.zeros(d)
    ctx.check(_capi.lib().cb2_kde_bw_loo(ctx._h, ptr(pts), d, N, ptr(circ), ptr(out)))
    return out


def kde_bandwidth_batch(pts_list, manifold=None, ctx=None):
    ctx = ctx or context()
    arrs = [f64(np.atleast_2d(p)) for p in pts_list]
    d = arrs[0].shape[1]
    B = len(arrs)
    circ = circ_arr(_topology(manifold, d), d)
    mus = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    Ns = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    out = np.zeros((B, d))
    ctx.check(_capi.lib().cb2_kde_bw_loo_batch(ctx._h, C.cast(mus, C.c_void_p), ptr(Ns), B, d, ptr(circ), ptr(out)))
    return out


def manikde(manifold, pts, bw=None, weights=None, partial=None, infoPerCoord=None, ctx=None):
    """`manikde!(M, pts; bw)`: bandwidth by LOO cross-validation on the device unless given."""
    pts = f64(np.atleast_2d(pts))
    if bw is None:
        bw = kde_bandwidth(pts, manifold, ctx=ctx)
    return ManifoldKernelDensity(manifold, pts, bw, weights, partial, infoPerCoord)


def mmd_manifold(a, b, circular, bw, rot_weight=2.0, ctx=None):
    """mmd with the product-manifold distance: wrapped differences on circular coordinates, weighted by `rot_weight`
    (2 = Manifolds.jl's metric on SpecialEuclidean(2): d_SO(2) = sqrt(2) |dtheta|)."""
    ctx = ctx or context()
    a, b = f64(np.atleast_2d(a)), f64(np.atleast_2d(b))
    circ = circ_arr(circular, a.shape[1])
    out = C.c_double()
    ctx.check(_capi.lib().cb2_mmd_manifold(ctx._h, ptr(a), a.shape[0], ptr(b), b.shape[0], a.shape[1], ptr(circ),
                                           float(rot_weight), float(np.ravel(bw)[0]), C.byref(out)))
    return out.value


def mmd(a, b, bw=None, rot_weight=2.0, ctx=None):
    """`mmd(M, a, b, N, M; bw)` on point arrays, or `mmd(mkdA, mkdB; bw=[0.001])` on two beliefs.
    Biased V-statistic MMD^2 with k = exp(-bw |p - q|^2).  Beliefs over variable types with circular
    coordinates (Pose2, Pose3) use the manifold distance (`mmd_manifold`)."""
    ctx = ctx or context()
    if isinstance(a, ManifoldKernelDensity):
        bw = 0.001 if bw is None else bw
        if any(a.circular):  # beliefs on SE(n): manifold distance (wrapped angles, product metric)
            return mmd_manifold(a.pts, b.pts, a.circular, bw, rot_weight=rot_weight, ctx=ctx)
        a, b = a.pts, b.pts
    if bw is None:
        raise ValueError("bw is required for point arrays")
    bw = float(np.ravel(bw)[0])
    a, b = f64(np.atleast_2d(a)), f64(np.atleast_2d(b))
    if a.shape[1] != b.shape[1]:
        raise ValueError("dimension mismatch")
    out = C.c_double()
    ctx.check(_capi.lib().cb2_mmd(ctx._h, ptr(a), a.shape[0], ptr(b), b.shape[0], a.shape[1], bw, C.byref(out)))
    return out.value


def mmd_sums(a, b, bw, rows_a=None, rows_b=None, ctx=None):
    """Raw pair sums {aa, ab, bb} over row ranges (the unit a row-sharded multi-GPU mmd all-reduces)."""
    ctx = ctx or context()
    a, b = f64(np.atleast_2d(a)), f64(np.atleast_2d(b))
    ra = rows_a or (0, a.shape[0])
    rb = rows_b or (0, b.shape[0])
    out = np.zeros(3)
    ctx.check(_capi.lib().cb2_mmd_sums(ctx._h, ptr(a), a.shape[0], ptr(b), b.shape[0], a.shape[1], float(bw),
                                       ra[0], ra[1], rb[0], rb[1], ptr(out)))
    return out


def mmd_sums_shard(a, b, bw, shard, nshards, ctx=None):
    """Shard `shard` of `nshards` of the three pair sums (tile list of the symmetric evaluation dealt round-robin): what one GPU
    of a multi-GPU mmd computes before the 3-value all-reduce."""
    ctx = ctx or context()
    a, b = f64(np.atleast_2d(a)), f64(np.atleast_2d(b))
    out = np.zeros(3)
    ctx.check(_capi.lib().cb2_mmd_sums_shard(ctx._h, ptr(a), a.shape[0], ptr(b), b.shape[0], a.shape[1], float(bw),
                                             int(shard), int(nshards), ptr(out)))
    return out


def mmd_combine(sums, N, M):
    return sums[0] / (N * N) - 2.0 * sums[1] / (N * M) + sums[2] / (M * M)


def isapprox(a, b, mmd_tol=1e-2, ctx=None):
    return mmd(a, b, ctx=ctx) < mmd_tol


def mmd_se_batch(a, b, bw, coords, backward=True, grad=True, ctx=None):
    """mmd(a, T(c_k) b) (or T(c_k)^-1 b when backward) for K coordinate vectors, with analytic gradients."""
    ctx = ctx or context()
    a, b = f64(np.atleast_2d(a)), f64(np.atleast_2d(b))
    c = f64(np.atleast_2d(coords))
    K, nc = c.shape
    dim = a.shape[1]
    if nc != (3 if dim == 2 else 6):
        raise ValueError(f"coords must have {3 if dim == 2 else 6} columns for dim={dim}")
    vals = np.zeros(K)
    grads = np.zeros((K, nc)) if grad else None
    ctx.check(_capi.lib().cb2_mmd_se_batch(ctx._h, ptr(a), a.shape[0], ptr(b), b.shape[0], dim, float(bw), ptr(c), K,
                                           int(backward), ptr(vals), ptr(grads)))
    return (vals, grads) if grad else vals


def evaluate(mkd, q, logp=False, ctx=None):
    ctx = ctx or context()
    q = f64(np.atleast_2d(q))
    if q.shape[1] != mkd.d:
        if q.shape[0] == mkd.d:  # Julia passes d x Q
            q = f64(q.T)
        else:
            raise ValueError("query dimension mismatch")
    out = np.zeros(q.shape[0])
    w = mkd.weights
    ctx.check(_capi.lib().cb2_kde_eval(ctx._h, ptr(mkd.pts), ptr(mkd.bw), ptr(w), mkd.d, mkd.N, ptr(q), q.shape[0],
                                       int(logp), ptr(out)))
    return out


def sample(mkd, n=1, seed=0, stream=0, ctx=None):
    """`sample(mkd, n)` -> (points, ): kernel ~ weights, plus bandwidth-scaled Gaussian noise."""
    ctx = ctx or context()
    out = np.zeros((int(n), mkd.d))
    circ = circ_arr(mkd.circular, mkd.d)
    ctx.check(_capi.lib().cb2_kde_sample(ctx._h, ptr(mkd.pts), ptr(mkd.bw), ptr(mkd.weights), mkd.d, mkd.N, ptr(circ),
                                         int(n), int(seed), int(stream), ptr(out)))
    return (out,)


def _pack_products(groups, Nouts, niter, streams, sample_offsets, want_labels, want_bw, d):
    B = len(groups)
    descs = (ProductDesc * B)()
    keep = []
    outs, labels, bws = [], [], []
    for b, dens in enumerate(groups):
        D = len(dens)
        arr = (Density * D)(*[m._density() for m in dens])
        keep.append(arr)
        o = np.zeros((Nouts[b], d))
        lab = np.zeros((Nouts[b], D), dtype=np.int32) if want_labels else None
        bw = np.zeros(d) if want_bw else None
        outs.append(o); labels.append(lab); bws.append(bw)
        descs[b].dens = arr
        descs[b].D = D
        descs[b].Nout = int(Nouts[b])
        descs[b].niter = int(niter)
        descs[b].stream = int(streams[b])
        descs[b].sample_offset = int(sample_offsets[b])
        descs[b].pts_out = o.ctypes.data
        descs[b].labels_out = lab.ctypes.data if lab is not None else None
        descs[b].bw_out = bw.ctypes.data if bw is not None else None
    return descs, keep, outs, labels, bws


def manifoldProducts(groups, manifold=None, N=None, Niter=1, seed=0, streams=None, sample_offsets=None,
                     return_labels=False, bw="loo", ctx=None):
    """Batched `manifoldProduct`: one device pass over B independent products (all variables of a Gibbs sweep /
    all ready cliques).  groups: list of lists of ManifoldKernelDensity over the same variable type.
    Returns a list of ManifoldKernelDensity (bw='loo': fresh LOO bandwidth as upstream; bw=None: skip)."""
    ctx = ctx or context()
    B = len(groups)
    if B == 0:
        return []
    first = groups[0][0]
    d = first.d
    manifold = manifold if manifold is not None else first.manifold
    topo = _topology(manifold, d)
    for g in groups:
        for m in g:
            if m.d != d or tuple(m.circular) != tuple(topo):
                raise ValueError("all messages of a batch must share the variable type")
    Nouts = [int(N) if N is not None else max(m.N for m in g) for g in groups]
    streams = list(range(B)) if streams is None else streams
    sample_offsets = [0] * B if sample_offsets is None else sample_offsets
    # upstream early-out: a product of one density is that density
    results = [None] * B
    multi = [b for b in range(B) if len(groups[b]) > 1]
    for b in range(B):
        if len(groups[b]) == 1:
            results[b] = (groups[b][0], None)
    if multi:
        want_bw = bw == "loo"
        descs, keep, outs, labels, bws = _pack_products([groups[b] for b in multi], [Nouts[b] for b in multi], Niter,
                                                        [streams[b] for b in multi], [sample_offsets[b] for b in multi],
                                                        return_labels, want_bw, d)
        circ = circ_arr(topo, d)
        ctx.check(_capi.lib().cb2_product_batch(ctx._h, descs, len(multi), ptr(circ), d, int(seed)))
        for i, b in enumerate(multi):
            bwv = bws[i] if want_bw else (bw if bw is not None else np.full(d, np.nan))
            results[b] = (ManifoldKernelDensity(manifold, outs[i], bwv), labels[i])
    if return_labels:
        return results
    return [r[0] for r in results]


def manifoldProduct(ff, manifold=None, Niter=1, N=None, seed=0, stream=0, return_labels=False, bw="loo", ctx=None):
    """`manifoldProduct(ff, M; Niter=1, N)`: Gibbs-sampled product of the beliefs in `ff`.  Inputs that carry a
    `partial` (informative coordinates only, 1-based as in the packed format) are multiplied dimension-wise, as
    upstream does."""
    ff = list(ff)
    if len(ff) > 1 and any(m.partial is not None for m in ff):
        return _manifoldProductPartial(ff, manifold, Niter, N, seed, stream, ctx)
    r = manifoldProducts([ff], manifold, N, Niter, seed, [stream], None, return_labels, bw, ctx)
    return r[0]


def _manifoldProductPartial(ff, manifold, Niter, N, seed, stream, ctx):
    """Dimension-wise product: coordinate i is the 1-D product of the marginals of every input that informs it."""
    first = ff[0]
    d = first.d
    manifold = manifold if manifold is not None else first.manifold
    topo = _topology(manifold, d)
    N = int(N) if N is not None else max(m.N for m in ff)
    cols, bws = [None] * d, np.zeros(d)
    batches = {0: [], 1: []}  # 1-D products grouped by topology (one circular mask per batched call)
    for i in range(d):
        use = [m for m in ff if m.partial is None or (i + 1) in tuple(m.partial)]
        if not use:
            use = [first]
        marg = [ManifoldKernelDensity((topo[i],), m.pts[:, i:i + 1], m.bw[i:i + 1], m.weights) for m in use]
        if len(marg) == 1:  # a single informed input: N fresh draws from that marginal (device sampler, Philox: reproducible)
            cols[i] = sample(marg[0], N, seed=seed, stream=stream * 16 + i, ctx=ctx)[0][:, 0]
            bws[i] = marg[0].bw[0]
        else:
            batches[topo[i]].append((i, marg))
    for c, items in batches.items():
        if not items:
            continue
        res = manifoldProducts([m for _, m in items], (c,), N=N, Niter=Niter, seed=seed,
                               streams=[stream * 16 + i for i, _ in items], bw="loo", ctx=ctx)
        for (i, _), r in zip(items, res):
            cols[i], bws[i] = r.pts[:, 0], r.bw[0]
    return ManifoldKernelDensity(manifold, np.stack(cols, axis=1), bws)


class PendingProducts:
    """Ticket of an asynchronous batch (`cb2_product_batch_submit`): the call returned as soon as the work was
    enqueued; `wait()` blocks until the results are in the host buffers (`cb2_wait`)."""

    def __init__(self, ctx, ticket, groups, manifold, outs, labels, bws, keep):
        self.ctx, self.ticket, self._groups, self._manifold = ctx, ticket, groups, manifold
        self._outs, self._labels, self._bws, self._keep = outs, labels, bws, keep

    def wait(self):
        self.ctx.check(_capi.lib().cb2_wait(self.ctx._h, self.ticket))
        return [ManifoldKernelDensity(self._manifold, o, b) for o, b in zip(self._outs, self._bws)]


def manifoldProducts_submit(groups, manifold=None, N=None, Niter=1, seed=0, streams=None, ctx=None):
    """Asynchronous flavour of `manifoldProducts` for cooperative callers (one Julia Task per clique): every group
    must hold at least two messages (the single-input early-out needs no device work)."""
    ctx = ctx or context()
    first = groups[0][0]
    d = first.d
    manifold = manifold if manifold is not None else first.manifold
    topo = _topology(manifold, d)
    if any(len(g) < 2 for g in groups):
        raise ValueError("submit: every product needs at least two messages")
    B = len(groups)
    Nouts = [int(N) if N is not None else max(m.N for m in g) for g in groups]
    streams = list(range(B)) if streams is None else streams
    descs, keep, outs, labels, bws = _pack_products(groups, Nouts, Niter, streams, [0] * B, False, True, d)
    circ = circ_arr(topo, d)
    ticket = C.c_int64()
    ctx.check(_capi.lib().cb2_product_batch_submit(ctx._h, descs, B, ptr(circ), d, int(seed), C.byref(ticket)))
    return PendingProducts(ctx, ticket.value, groups, manifold, outs, labels, bws, (keep, descs, circ))


def product_points(dens, circular, Nout, seed=0, niter=1, stream=0, sample_offset=0, ctx=None):
    """Raw single product: dens = list of (pts, bw[, w]).  Returns (pts Nout x d, labels Nout x D)."""
    ctx = ctx or context()
    ms = [ManifoldKernelDensity(circular, x[0], x[1], x[2] if len(x) > 2 else None) for x in dens]
    d = ms[0].d
    descs, keep, outs, labels, _ = _pack_products([ms], [Nout], niter, [stream], [sample_offset], True, False, d)
    circ = circ_arr(ms[0].circular, d)
    ctx.check(_capi.lib().cb2_product_batch(ctx._h, descs, 1, ptr(circ), d, int(seed)))
    return outs[0], labels[0]


FACTOR_PRIOR_POSE2, FACTOR_POSE2POSE2, FACTOR_POSE2POINT2_BEARINGRANGE, FACTOR_PRIOR_POSE3, FACTOR_POSE3POSE3 = 0, 1, 2, 3, 4
FACTOR_PRIOR_POINT2, FACTOR_POINT2POINT2_RANGE, FACTOR_LINEAR_SAMPLES = 5, 6, 7


def _conv_dims(kind, rev, mean=None):
    """(coordinates of the noise / mean, coordinates of the target) of a closed-form convolution."""
    if kind == FACTOR_LINEAR_SAMPLES:
        return 3, int(np.asarray(mean, dtype=np.float64).ravel()[0])
    if kind in (FACTOR_PRIOR_POSE3, FACTOR_POSE3POSE3):
        return 6, 6
    if kind in (FACTOR_PRIOR_POINT2, FACTOR_POINT2POINT2_RANGE):
        return 3, 2
    return 3, (2 if (kind == FACTOR_POSE2POINT2_BEARINGRANGE and not rev) else 3)


def approxConvBatch(convs, seed=0, ctx=None):
    """Closed-form `approxConvBelief` for a batch of factors in one device launch.
    convs: list of dicts {kind, reverse, N, mean(3), chol(3x3 lower), stream, src (n x d) | None, aux (n x 3) | None}.
    Returns the list of N x d_tgt sample arrays."""
    ctx = ctx or context()
    B = len(convs)
    descs = (ConvDesc * B)()
    keep, outs = [], []
    for b, c in enumerate(convs):
        kind, rev = int(c["kind"]), int(bool(c.get("reverse", False)))
        nz, dt = _conv_dims(kind, rev, c.get("mean"))
        out = np.zeros((int(c["N"]), dt))
        src = None if c.get("src") is None else f64(c["src"])
        aux = None if c.get("aux") is None else f64(c["aux"])
        keep += [src, aux]
        outs.append(out)
        d = descs[b]
        d.kind, d.reverse, d.N, d.stream = kind, rev, int(c["N"]), int(c.get("stream", b))
        d.n_src = 0 if src is None else src.shape[0]
        d.n_aux = 0 if aux is None else aux.shape[0]
        d.src = None if src is None else src.ctypes.data
        d.aux = None if aux is None else aux.ctypes.data
        d.mean[:nz] = np.resize(np.asarray(c["mean"], dtype=np.float64), nz).tolist()
        d.chol[:nz * nz] = np.asarray(c["chol"], dtype=np.float64).reshape(nz * nz).tolist()
        d.out = out.ctypes.data
    ctx.check(_capi.lib().cb2_approx_conv_batch(ctx._h, descs, B, int(seed)))
    return outs


def se_residual_batch(dim, X, p, q, ctx=None):
    """Residual of the relative-pose factors (ScatterAlignPose2/3, Pose2Pose2, Pose3Pose3) for K particles in one launch:
    `Xc = vee(M, q, log(M, q, p o exp(M, e0, X)))` (R:ext/Images/ScatterAlignPose2.jl:216-231).  X, p, q: K x ncoord coordinates."""
    ctx = ctx or context()
    X, p, q = (f64(np.atleast_2d(v)) for v in (X, p, q))
    nc = 3 if dim == 2 else 6
    if X.shape != p.shape or X.shape != q.shape or X.shape[1] != nc:
        raise ValueError(f"X, p, q must all be K x {nc}")
    out = np.zeros_like(X)
    ctx.check(_capi.lib().cb2_se_residual_batch(ctx._h, int(dim), ptr(X), ptr(p), ptr(q), X.shape[0], ptr(out)))
    return out


def icp_align_batch(fix, mov, perturb=None, correspondences=500, neighbors=50, min_planarity=0.3, min_change=3.0,
                    max_iterations=25, ctx=None):
    """`_PCL.alignICP_Simple(X_fix, X_mov; ...)` (R:src/3rdParty/_PCL/services/ICP_Simple.jl:239) for K perturbed copies of
    `X_mov` in one device batch.  Returns (H [K, 4, 4] fix_H_mov, stats [K, 4] = mean / std of the last residuals,
    correspondences kept, iterations)."""
    ctx = ctx or context()
    fix, mov = f64(fix), f64(mov)
    if fix.ndim != 2 or fix.shape[1] != 3:
        raise ValueError('"X_fix" must have 3 columns')
    if mov.ndim != 2 or mov.shape[1] != 3:
        raise ValueError('"X_mov" must have 3 columns')
    pert = None if perturb is None else f64(np.atleast_2d(perturb))
    K = 1 if pert is None else pert.shape[0]
    H, st = np.zeros((K, 4, 4)), np.zeros((K, 4))
    ctx.check(_capi.lib().cb2_icp_align_batch(ctx._h, ptr(fix), fix.shape[0], ptr(mov), mov.shape[0], ptr(pert), K,
                                              int(correspondences), int(neighbors), float(min_planarity), float(min_change),
                                              int(max_iterations), ptr(H), ptr(st)))
    return H, st


def tree_levels(pts, bw, w=None, ctx=None):
    """Level hierarchy of one message as built on the device (parity tests)."""
    ctx = ctx or context()
    m = ManifoldKernelDensity(None, pts, bw, w)
    den = m._density()
    depth, cnt = C.c_int32(), C.c_int32()
    perm = np.zeros(m.N, dtype=np.int32)
    L = _capi.lib()
    ctx.check(L.cb2_tree_debug(ctx._h, C.byref(den), m.d, 0, C.byref(depth), C.byref(cnt), ptr(perm), None, None, None))
    levels = []
    for l in range(depth.value + 1):
        ctx.check(L.cb2_tree_debug(ctx._h, C.byref(den), m.d, l, C.byref(depth), C.byref(cnt), None, None, None, None))
        mean, var, wt = np.zeros((cnt.value, m.d)), np.zeros((cnt.value, m.d)), np.zeros(cnt.value)
        ctx.check(L.cb2_tree_debug(ctx._h, C.byref(den), m.d, l, C.byref(depth), C.byref(cnt), None, ptr(mean), ptr(var), ptr(wt)))
        levels.append((mean, var, wt))
    return depth.value, perm, levels


from .factors import (ScatterAlign, ScatterAlignPose2, ScatterAlignPose3, CalcFactor, preambleCache, getSample,  # noqa: E402
                      HeatmapGridDensity, ScatterAlignPose2_from_images, getManifold, parchDistribution, packScatterAlign,
                      unpackScatterAlign, BlobStoreGraph)
from .serialization import packDistribution, unpackDistribution, multiplyDistributions, solveTree  # noqa: E402
