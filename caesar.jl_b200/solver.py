"""Caller harness for the batched product path: a small nonparametric Gibbs solver over Pose2 / Point2 graphs.

This is NOT a port of IncrementalInference's Bayes-tree / clique-state-machine solver (out of scope).  It
plays the role IIF's `propagateBelief` plays for the hot path -- per variable: one proposal per factor
(`approxConvBelief`, closed form for the group factors used here: SURVEY sec. 8f N1), `manikde!` of every proposal,
`manifoldProduct`, `manikde!` of the result -- and issues those products the way the Julia shim would: variables that
share no factor are batched into ONE cb2_product_batch call per Gibbs colour (K6), proposals' LOO bandwidths in one
cb2_kde_bw_loo_batch call.  Used for BASELINE config C2 (hexagonal Pose2 SLAM, `generateGraph_Hexagonal`;
recipe R:docs/src/examples/basic_hexagonal2d.md:21-43,89-111) and as an end-to-end parity test of the product path.

The whole graph is treated as one clique (what IIF does inside the root clique): `gibbs_iters` sweeps (IIF default 3)
after a forward initialisation (`initAll!`).
"""
import time
from dataclasses import dataclass, field

import numpy as np

TWO_PI = 2.0 * np.pi
VARTYPES = ("Pose2", "Point2", "ContinuousScalar")


def _wrap(a):
    return (a + np.pi) % TWO_PI - np.pi


@dataclass
class Variable:
    label: str
    vartype: str  # "Pose2" | "Point2"
    belief: object = None  # ManifoldKernelDensity


@dataclass
class PriorPose2:
    mean: np.ndarray
    cov: np.ndarray


@dataclass
class Pose2Pose2:
    """`multihypo = [1, p_1, .., p_H]` over labels [x, h_1, .., h_H] (IIF `addFactor!(...; multihypo)`,
    R:test/ICRA2022_tests/multihypo_corners.jl:55-60): the second variable of the relative factor is h_k with probability p_k, chosen
    per particle.  `nullhypo = p`: with probability p a particle treats the factor as uninformative (:103)."""
    mean: np.ndarray  # [dx, dy, dtheta] in the first pose's frame
    cov: np.ndarray
    multihypo: object = None
    nullhypo: float = 0.0


@dataclass
class Pose2Point2BearingRange:
    bearing: tuple  # (mean, sigma)
    range: tuple


@dataclass
class PriorPoint2:
    mean: np.ndarray
    cov: np.ndarray


@dataclass
class Point2Point2Range:
    """RoME.Point2Point2Range(Normal(rho, sigma)): one constraint on two coordinates.  The proposal is the ring of radius ~ N(rho, sigma)
    around the other variable's particles with a uniform direction (what IIF's per-particle solve produces from scattered starting
    points; R:test/ICRA2022_tests/insufficient_data.jl:70-85,126-142)."""
    rho: float
    sigma: float


# ---- scalar factors with arbitrary noise (IIF.Prior / LinearRelative / Mixture on ContinuousScalar: the graph of the ICRA tutorial 2
# test R:test/ICRA2022_tests/nongaussian_mixture.jl:18-51).  A distribution is a sampler `(rng, N) -> N values`; the factor's own
# `getSample` stays on the host (as upstream's stays in Julia), the solve x_j = x_i + z is convolution kind 7 of the library.
def Normal(mu, sigma):
    return lambda rng, N: mu + sigma * rng.standard_normal(N)


def Rayleigh(sigma):
    return lambda rng, N: rng.rayleigh(sigma, size=N)


def Uniform(a, b):
    return lambda rng, N: rng.uniform(a, b, size=N)


def MixtureDist(components, weights):
    """`Mixture(LinearRelative, (a = .., b = ..), [w_a; w_b])`: every particle draws its component with the given probabilities."""
    w = np.asarray(weights, dtype=np.float64)

    def draw(rng, N):
        pick = rng.choice(len(w), size=N, p=w / w.sum())
        out = np.empty(N)
        for k, c in enumerate(components):
            sel = np.flatnonzero(pick == k)
            out[sel] = c(rng, len(sel))
        return out
    return draw


@dataclass
class Prior:
    dist: object


@dataclass
class LinearRelative:
    dist: object


@dataclass
class FactorGraph:
    variables: dict = field(default_factory=dict)
    factors: list = field(default_factory=list)  # (labels, factor)
    N: int = 100

    def addVariable(self, label, vartype):
        self.variables[label] = Variable(label, vartype)

    def addFactor(self, labels, factor):
        self.factors.append((list(labels), factor))

    def factors_of(self, label):
        return [(ls, f) for ls, f in self.factors if label in ls]


def generateGraph_Hexagonal(N=100, landmark=True, loop_closure=True):
    """RoME.generateGraph_Hexagonal as spelled out in R:docs/src/examples/basic_hexagonal2d.md."""
    fg = FactorGraph(N=N)
    fg.addVariable("x0", "Pose2")
    fg.addFactor(["x0"], PriorPose2(np.zeros(3), 0.01 * np.eye(3)))
    for i in range(6):
        fg.addVariable(f"x{i + 1}", "Pose2")
        fg.addFactor([f"x{i}", f"x{i + 1}"], Pose2Pose2(np.array([10.0, 0.0, np.pi / 3]), np.diag([0.1, 0.1, 0.1]) ** 2))
    if landmark:
        fg.addVariable("l1", "Point2")
        fg.addFactor(["x0", "l1"], Pose2Point2BearingRange((0.0, 0.1), (20.0, 1.0)))
        if loop_closure:
            fg.addFactor(["x6", "l1"], Pose2Point2BearingRange((0.0, 0.1), (20.0, 1.0)))
    return fg


# ---------------------------------------------------------------- closed-form convolutions (approxConvBelief)
def _compose(p, X):
    """p o exp_e(X) on SE(2) with the product exponential (translation = coordinates)."""
    c, s = np.cos(p[:, 2]), np.sin(p[:, 2])
    return np.c_[p[:, 0] + c * X[:, 0] - s * X[:, 1], p[:, 1] + s * X[:, 0] + c * X[:, 1], _wrap(p[:, 2] + X[:, 2])]


def _compose_inv(q, X):
    """q o exp_e(X)^-1"""
    th = _wrap(q[:, 2] - X[:, 2])
    c, s = np.cos(th), np.sin(th)
    return np.c_[q[:, 0] - (c * X[:, 0] - s * X[:, 1]), q[:, 1] - (s * X[:, 0] + c * X[:, 1]), th]


def _resample(pts, N, rng):
    return pts if len(pts) == N else pts[rng.integers(0, len(pts), size=N)]


def approx_conv(fg, labels, factor, target, rng):
    """Samples of `target` implied by `factor` and the current belief of the other variable (N points)."""
    N = fg.N
    if isinstance(factor, PriorPose2):
        x = rng.multivariate_normal(factor.mean, factor.cov, size=N)
        x[:, 2] = _wrap(x[:, 2])
        return x
    if isinstance(factor, PriorPoint2):
        return rng.multivariate_normal(factor.mean, factor.cov, size=N)
    if isinstance(factor, Prior):
        return factor.dist(rng, N)[:, None]
    other = labels[0] if labels[1] == target else labels[1]
    src = _resample(fg.variables[other].belief.pts, N, rng)
    if isinstance(factor, LinearRelative):
        z = factor.dist(rng, N)[:, None]
        return src + z if target == labels[1] else src - z
    if isinstance(factor, Point2Point2Range):
        r = factor.rho + factor.sigma * rng.standard_normal(N)
        th = rng.uniform(-np.pi, np.pi, size=N)
        return np.c_[src[:, 0] + r * np.cos(th), src[:, 1] + r * np.sin(th)]
    if isinstance(factor, Pose2Pose2):
        X = rng.multivariate_normal(factor.mean, factor.cov, size=N)
        return _compose(src, X) if target == labels[1] else _compose_inv(src, X)
    if isinstance(factor, Pose2Point2BearingRange):
        b = factor.bearing[0] + factor.bearing[1] * rng.standard_normal(N)
        r = factor.range[0] + factor.range[1] * rng.standard_normal(N)
        if target == labels[1]:  # pose -> landmark
            return np.c_[src[:, 0] + r * np.cos(src[:, 2] + b), src[:, 1] + r * np.sin(src[:, 2] + b)]
        # landmark -> pose: two constraints on three coordinates; the heading comes from the pose's current belief
        th = _resample(fg.variables[target].belief.pts, N, rng)[:, 2]
        return np.c_[src[:, 0] - r * np.cos(th + b), src[:, 1] - r * np.sin(th + b), th]
    raise TypeError(f"unsupported factor {type(factor).__name__}")


SPREAD_NH = 3.0  # IIF solver default `spreadNH`: how far (in factor sigmas) null-hypothesis particles are scattered


def _host_pts(fg, label):
    b = fg.variables[label].belief
    if hasattr(b, "to_host"):
        raise NotImplementedError("multihypo / nullhypo factors mix particles on the host: use a host-pointer backend")
    return b.pts


def _hypo_desc(fg, labels, factor, target, stream, rng):
    """Pose2Pose2 with fractional associations.  The association is drawn per particle on the host (same
    generator for every backend); the convolution itself stays one closed-form device launch.  Returns the descriptor plus a
    `post` rule (keep_mask, own_points, scatter): particles under `keep_mask` do not take the factor's proposal but one of the
    target's current points, scattered by `scatter` sigmas of the factor noise when the particle chose the null hypothesis."""
    N = fg.N
    chol = np.linalg.cholesky(factor.cov)
    own = fg.variables[target].belief
    own_pts = None if own is None else _resample(_host_pts(fg, target), N, rng)
    keep = np.zeros(N, dtype=bool)
    scatter = np.zeros(N)
    if factor.multihypo is not None:
        p = np.asarray(factor.multihypo[1:], dtype=np.float64)
        hyp = rng.choice(len(p), size=N, p=p / p.sum())
        if target == labels[0]:  # the certain variable: every particle is solved from the hypothesis it drew
            src = np.empty((N, 3))
            for k in range(len(p)):
                sel = np.flatnonzero(hyp == k)
                pk = _host_pts(fg, labels[1 + k])
                src[sel] = pk[rng.integers(0, len(pk), size=len(sel))]
            reverse = True
        else:  # hypothesis variable k: only the particles that drew k are informed by the factor
            k = labels.index(target) - 1
            src, reverse = _host_pts(fg, labels[0]), False
            keep = hyp != k
    else:
        other = labels[0] if labels[1] == target else labels[1]
        src, reverse = _host_pts(fg, other), target == labels[0]
    if factor.nullhypo > 0 and own_pts is not None:
        null = rng.uniform(size=N) < factor.nullhypo
        scatter[null & ~keep] = SPREAD_NH
        keep = keep | null
    desc = dict(kind=1, reverse=reverse, N=N, mean=factor.mean, chol=chol, stream=stream, src=src, aux=None)
    if own_pts is not None and keep.any():
        desc["post"] = (keep, own_pts, scatter, np.sqrt(np.diag(factor.cov)), rng.standard_normal((N, 3)))
    return desc


def _apply_post(out, post):
    keep, own, scatter, sig, z = post
    out = np.array(out, copy=True)
    rep = own + (scatter[:, None] * sig[None, :]) * z
    rep[:, 2] = _wrap(rep[:, 2])
    out[keep] = rep[keep]
    return out


def _chol(factor):
    """Cholesky factor of the factor's noise covariance, computed once per factor object."""
    L = factor.__dict__.get("_chol")
    if L is None:
        L = np.linalg.cholesky(factor.cov)
        factor.__dict__["_chol"] = L
    return L


def conv_desc(fg, labels, factor, target, stream, rng=None):
    """Descriptor of one closed-form convolution for a backend's `approx_conv_batch` (device: `approxConvBatch`)."""
    N = fg.N
    if isinstance(factor, PriorPose2):
        return dict(kind=0, reverse=False, N=N, mean=factor.mean, chol=_chol(factor), stream=stream, src=None, aux=None)
    if isinstance(factor, PriorPoint2):
        L = np.zeros((3, 3))
        L[:2, :2] = np.linalg.cholesky(factor.cov)
        return dict(kind=5, reverse=False, N=N, mean=[factor.mean[0], factor.mean[1], 0.0], chol=L, stream=stream, src=None, aux=None)
    if isinstance(factor, Prior):
        return dict(kind=7, reverse=False, N=N, mean=[1.0, 0.0, 0.0], chol=np.zeros((3, 3)), stream=stream, src=None,
                    aux=factor.dist(rng, N)[:, None])
    if isinstance(factor, Pose2Pose2) and (factor.multihypo is not None or factor.nullhypo > 0):
        return _hypo_desc(fg, labels, factor, target, stream, rng)
    other = labels[0] if labels[1] == target else labels[1]
    src = fg.variables[other].belief.pts
    if isinstance(factor, LinearRelative):
        return dict(kind=7, reverse=target == labels[0], N=N, mean=[1.0, 0.0, 0.0], chol=np.zeros((3, 3)), stream=stream, src=src,
                    aux=factor.dist(rng, N)[:, None])
    if isinstance(factor, Point2Point2Range):
        return dict(kind=6, reverse=target == labels[0], N=N, mean=[factor.rho, 0.0, 0.0], chol=np.diag([factor.sigma, 0.0, 0.0]),
                    stream=stream, src=src, aux=None)
    if isinstance(factor, Pose2Pose2):
        return dict(kind=1, reverse=target == labels[0], N=N, mean=factor.mean, chol=_chol(factor), stream=stream, src=src, aux=None)
    if isinstance(factor, Pose2Point2BearingRange):
        rev = target == labels[0]
        return dict(kind=2, reverse=rev, N=N, mean=[factor.bearing[0], factor.range[0], 0.0],
                    chol=np.diag([factor.bearing[1], factor.range[1], 0.0]), stream=stream, src=src,
                    aux=fg.variables[target].belief.pts if rev else None)
    raise TypeError(f"unsupported factor {type(factor).__name__}")


def _convolve(fg, backend, items, rng, seed, stream0):
    """items: [(labels, factor, target)].  Uses the backend's batched closed-form convolution when it has one
    (device: one cb2_approx_conv_batch launch), else the NumPy fallback of this harness."""
    if hasattr(backend, "approx_conv_batch"):
        descs = [conv_desc(fg, ls, f, t, stream0 + i, rng) for i, (ls, f, t) in enumerate(items)]
        posts = [d.pop("post", None) for d in descs]
        outs = backend.approx_conv_batch(descs, seed)
        return [o if p is None else _apply_post(o, p) for o, p in zip(outs, posts)]
    return [approx_conv(fg, ls, f, t, rng) for ls, f, t in items]


class _PointsOnly:
    """An initialised variable before its bandwidth has been searched (forward initialisation)."""

    def __init__(self, pts):
        self.pts = pts


def _colouring(fg):
    """Greedy colouring: variables of one colour share no factor and are updated in one batched product call."""
    colour, order = {}, list(fg.variables)
    for v in order:
        nb = {l for ls, _ in fg.factors_of(v) for l in ls if l != v}
        used = {colour[n] for n in nb if n in colour}
        colour[v] = next(c for c in range(len(order)) if c not in used)
    groups = {}
    for v, c in colour.items():
        groups.setdefault(c, []).append(v)
    return [groups[c] for c in sorted(groups)]


def solveGraph(fg, backend, gibbs_iters=3, seed=0, stats=None):
    """backend: object with manikde_batch(vartype, [pts]) -> [MKD] and products(vartype, [[MKD]], N, seed, streams) -> [MKD].
    Returns stats: wall seconds, number and shapes of products issued."""
    rng = np.random.default_rng(seed)
    stats = stats if stats is not None else {}
    stats.update(products=0, shapes={}, product_calls=0, t_products=0.0, t_bandwidth=0.0)
    t0 = time.perf_counter()
    if hasattr(backend, "begin"):
        backend.begin(fg)
    # forward initialisation (initAll!): a variable is initialised from the first factor whose other end already is
    pending = [v for v in fg.variables if fg.variables[v].belief is None]  # (a re-solve after adding to the graph keeps the rest)
    fresh = list(pending)
    while pending:
        progressed = False
        for v in list(pending):
            for ls, f in fg.factors_of(v):
                others = [l for l in ls if l != v]
                if all(fg.variables[o].belief is not None for o in others) and not (
                        isinstance(f, Pose2Point2BearingRange) and v == ls[0]):
                    pts = _convolve(fg, backend, [(ls, f, v)], rng, seed * 7919 + 1, 900000 + len(pending))[0]
                    fg.variables[v].belief = _PointsOnly(pts)   # the chain of initialisations needs points only
                    pending.remove(v)
                    progressed = True
                    break
        if not progressed:
            raise RuntimeError(f"cannot initialise {pending}")
    for vartype in VARTYPES:  # the bandwidths of all initial beliefs in one batched search per variable type
        vs = [v for v in fresh if fg.variables[v].vartype == vartype]
        if vs:
            for v, b in zip(vs, backend.manikde_batch(vartype, [fg.variables[v].belief.pts for v in vs])):
                fg.variables[v].belief = b
    colours = _colouring(fg)
    stream = 0
    conv_stream = 0
    for it in range(gibbs_iters):
        for group in colours:
            for vartype in VARTYPES:
                vs = [v for v in group if fg.variables[v].vartype == vartype]
                if not vs:
                    continue
                items, counts = [], []
                for v in vs:
                    fs = fg.factors_of(v)
                    items += [(ls, f, v) for ls, f in fs]
                    counts.append(len(fs))
                tc = time.perf_counter()
                props = _convolve(fg, backend, items, rng, seed * 7919 + 2 + it, conv_stream)
                conv_stream += len(items)
                stats["t_conv"] = stats.get("t_conv", 0.0) + time.perf_counter() - tc
                tb = time.perf_counter()
                mkds = backend.manikde_batch(vartype, props)
                stats["t_bandwidth"] += time.perf_counter() - tb
                groups, k = [], 0
                for c in counts:
                    groups.append(mkds[k:k + c])
                    k += c
                tp = time.perf_counter()
                res = backend.products(vartype, groups, fg.N, seed * 100003 + it, list(range(stream, stream + len(vs))))
                stats["t_products"] += time.perf_counter() - tp
                stream += len(vs)
                stats["product_calls"] += 1
                for v, g, r in zip(vs, groups, res):
                    fg.variables[v].belief = r
                    if len(g) > 1:
                        stats["products"] += 1
                        key = f"D={len(g)},N={fg.N},d={r.d}"
                        stats["shapes"][key] = stats["shapes"].get(key, 0) + 1
    if hasattr(backend, "finish"):  # device-resident beliefs: read everything back (the only point where the host waits)
        backend.finish(fg)
    stats["seconds"] = time.perf_counter() - t0
    return stats


def generateGraph_MultihypoCorners(N=1000, stage=3):
    """The multimodal Pose2 graph of R:test/ICRA2022_tests/multihypo_corners.jl:14-148 (BASELINE config C4), built in the
    reference's stages: 0 = four corner poses and two door poses with priors (:17-38); 1 = x0 seen from one of the four corners,
    multihypo = [1, .25, .25, .25, .25] (:53-60); 2 = x1 with a 50 % null-hypothesis odometry factor and its own corner sighting
    (:102-111); 3 = x2 (:143-146).  `extend_MultihypoCorners` adds the next stage to a solved graph, as the test does."""
    fg = FactorGraph(N=N)
    for st in range(stage + 1):
        extend_MultihypoCorners(fg, st)
    return fg


def extend_MultihypoCorners(fg, stage):
    prior = np.diag([0.1, 0.1, 0.01]) ** 2
    dual = np.diag([1.0, 1.0, np.sqrt(np.pi)]) ** 2
    odo = np.diag([0.5, 0.5, 0.05]) ** 2
    corners = ["c0", "c1", "c2", "c3"]
    mh = [1.0, 0.25, 0.25, 0.25, 0.25]
    if stage == 0:
        for lb, mean, cov in (("c0", [0.0, 0, 0], prior), ("c1", [20.0, 0, np.pi / 2], prior), ("c2", [20.0, 10, np.pi], prior),
                              ("c3", [0.0, 10, -np.pi / 2], prior), ("door_prior", [20.0, 4, 0], prior), ("door_dual", [20.0, 4, np.pi], dual)):
            fg.addVariable(lb, "Pose2")
            fg.addFactor([lb], PriorPose2(np.array(mean), cov))
    elif stage == 1:
        fg.addVariable("x0", "Pose2")
        fg.addFactor(["x0"] + corners, Pose2Pose2(np.array([-2.0, -2, 0]), odo, multihypo=mh))
    elif stage == 2:
        fg.addVariable("x1", "Pose2")
        fg.addFactor(["x0", "x1"], Pose2Pose2(np.array([4.0, 0, np.pi / 2]), odo, nullhypo=0.5))
        fg.addFactor(["x1"] + corners, Pose2Pose2(np.array([-2.0, -4, 0]), odo, multihypo=mh))
    elif stage == 3:
        fg.addVariable("x2", "Pose2")
        fg.addFactor(["x1", "x2"], Pose2Pose2(np.array([16.0, 0, 0]), odo))
        fg.addFactor(["x2"] + corners, Pose2Pose2(np.array([2.0, -4, np.pi / 2]), odo, multihypo=mh))
    else:
        raise ValueError("stage 0..3")


TUT3_POSES = {"x0": (0.0, 0.0), "x1": (50.0, 0.0), "x2": (100.0, 0.0), "x3": (100.0, 50.0), "x4": (100.0, 100.0), "x5": (50.0, 100.0),
              "x6": (0.0, 100.0), "x7": (0.0, 50.0), "x8": (0.0, -50.0)}
TUT3_BEACONS = {"l1": (10.0, 30.0), "l2": (30.0, -30.0), "l3": (80.0, 40.0), "l4": (120.0, -50.0)}


def generateGraph_Tut3(N=100):
    """Range-only SLAM of the ICRA 2022 tutorial 3, first stage (R:test/ICRA2022_tests/insufficient_data.jl:9-27,104-134): x0 with
    no prior, three beacons of which two carry priors, one range measurement to each."""
    fg = FactorGraph(N=N)
    for v in ("x0", "l1", "l2", "l3"):
        fg.addVariable(v, "Point2")
    for l in ("l1", "l2"):
        fg.addFactor([l], PriorPoint2(np.array(TUT3_BEACONS[l]), np.eye(2)))
    for l, sigma in (("l1", 2.0), ("l2", 3.0), ("l3", 3.0)):
        fg.addFactor(["x0", l], Point2Point2Range(float(np.hypot(*np.subtract(TUT3_BEACONS[l], TUT3_POSES["x0"]))), sigma))
    return fg


def tut3_vehicle_drives(fg, frm, to, measurelimit=150.0):
    """`vehicle_drives!` of the same test (:66-93): an odometry range to the new position, then a range to every beacon in reach."""
    if to not in fg.variables:
        fg.addVariable(to, "Point2")
        fg.addFactor([frm, to], Point2Point2Range(float(np.hypot(*np.subtract(TUT3_POSES[frm], TUT3_POSES[to]))), 3.0))
    for l, pos in TUT3_BEACONS.items():
        rho = float(np.hypot(*np.subtract(pos, TUT3_POSES[to])))
        if rho < measurelimit:
            if l not in fg.variables:
                fg.addVariable(l, "Point2")
            fg.addFactor([to, l], Point2Point2Range(rho, 3.0))


def getPPE(belief):
    """Point estimate: Euclidean mean, circular mean on wrapped coordinates."""
    if hasattr(belief, "to_host"):  # device-resident belief: this is where the host waits
        belief = belief.to_host()
    out = []
    for i, c in enumerate(belief.circular):
        x = belief.pts[:, i]
        out.append(np.arctan2(np.sin(x).mean(), np.cos(x).mean()) if c else x.mean())
    return np.array(out)


class DeviceBackend:
    """Products and bandwidths through libcaesar_b200.so (the path under test)."""

    def __init__(self, cb, ctx=None):
        self.cb, self.ctx = cb, ctx or cb.context()

    def manikde_batch(self, vartype, pts_list):
        bws = self.cb.kde_bandwidth_batch(pts_list, vartype, ctx=self.ctx)
        return [self.cb.ManifoldKernelDensity(vartype, p, b) for p, b in zip(pts_list, bws)]

    def products(self, vartype, groups, N, seed, streams):
        return self.cb.manifoldProducts(groups, vartype, N=N, Niter=1, seed=seed, streams=streams, bw="loo", ctx=self.ctx)

    def approx_conv_batch(self, convs, seed):
        return self.cb.approxConvBatch(convs, seed=seed, ctx=self.ctx)


class DeviceResidentBackend:
    """beliefs stay on the device across the whole solve.  Proposals (approxConvBelief), bandwidths
    (manikde!) and products are stream-ordered `*_dev` calls on device buffers carved from one slab; nothing waits for the
    host until a belief is read back (`getPPE` / `to_host`).  Same Philox counters as DeviceBackend: in FP64 mode both
    backends produce the same points."""

    def __init__(self, cb, ctx=None, slab_bytes=64 << 20):
        from . import resident
        self.cb, self.ctx, self.r = cb, ctx or cb.context(), resident
        self.slab = resident.DeviceSlab(self.ctx, slab_bytes)

    def manikde_batch(self, vartype, arrays):
        return self.r.manikde_dev(vartype, arrays, self.slab, self.ctx)

    def products(self, vartype, groups, N, seed, streams):
        return self.r.manifoldProducts_dev(groups, vartype, N, self.slab, Niter=1, seed=seed, streams=streams, ctx=self.ctx)

    def approx_conv_batch(self, convs, seed):
        for c in convs:  # noise samples the host drew (convolution kind 7) travel to the device first
            for k in ("src", "aux"):
                if isinstance(c.get(k), np.ndarray):
                    c[k] = self.r.upload_array(c[k], self.slab, self.ctx)
        return self.r.approxConvBatch_dev(convs, self.slab, seed=seed, ctx=self.ctx)

    def begin(self, fg):
        """A re-solve after the graph has grown starts from host beliefs: they go back to the device first."""
        for v in fg.variables.values():
            if v.belief is not None and not hasattr(v.belief, "to_host"):
                v.belief = self.r.upload_belief(v.belief, self.slab, self.ctx)

    def finish(self, fg):
        """Downloads every belief (the one synchronisation of the solve)."""
        vs = [v for v in fg.variables.values() if hasattr(v.belief, "to_host")]
        for v, b in zip(vs, self.r.download_beliefs([v.belief for v in vs], self.ctx)):
            v.belief = b
