"""ctypes loader for the CPU FP64 restatement (oracle/cb2_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Density(C.Structure):
    _fields_ = [("N", C.c_int32), ("pts", C.c_void_p), ("bw", C.c_void_p), ("w", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libcb2_oracle.so")
    src = os.path.join(_HERE, "cb2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.cb2o_uniform.restype = C.c_double
        _LIB.cb2o_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        _LIB.cb2o_normal.restype = C.c_double
        _LIB.cb2o_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        _LIB.cb2o_loo_nll_1d.restype = C.c_double
        _LIB.cb2o_loo_nll_1d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        _LIB.cb2o_tree_build.restype = C.c_void_p
        for f in ("perm", "level_mean", "level_var", "level_wt"):
            getattr(_LIB, "cb2o_tree_" + f).restype = C.c_void_p
    return _LIB


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _circ(circular, d):
    if circular is None:
        return None
    c = np.ascontiguousarray(circular, dtype=np.uint8)
    assert c.size == d
    return c


def num_threads():
    return lib().cb2o_num_threads()


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs ask for every host core explicitly."""
    lib().cb2o_set_num_threads(len(os.sched_getaffinity(0)))
    return num_threads()


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().cb2o_philox4x32(_p(c), _p(k), _p(o))
    return o


def uniform(seed, sample, draw, stream, tag, ubits=53):
    return lib().cb2o_uniform(seed, sample, draw, stream, tag, ubits)


def mmd(a, b, bw, return_sums=False):
    """a: N x d, b: M x d (row = point)."""
    a, b = _f64(a), _f64(b)
    out = C.c_double()
    sums = np.zeros(3)
    rc = lib().cb2o_mmd(_p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], C.c_double(bw), C.byref(out), _p(sums))
    assert rc == 0
    return (out.value, sums) if return_sums else out.value


def mmd_rows(a, b, bw, ra, rb):
    a, b = _f64(a), _f64(b)
    sums = np.zeros(3)
    rc = lib().cb2o_mmd_rows(_p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], C.c_double(bw),
                             ra[0], ra[1], rb[0], rb[1], _p(sums))
    assert rc == 0
    return sums


def se_matrix(dim, coords, backward=False):
    c = _f64(coords)
    R = np.zeros((dim, dim))
    t = np.zeros(dim)
    lib().cb2o_se_matrix(dim, _p(c), int(backward), _p(R), _p(t))
    return R, t


def transform_points(dim, coords, pts, backward=False):
    c, p = _f64(coords), _f64(pts)
    out = np.zeros_like(p)
    lib().cb2o_transform_points(dim, _p(c), int(backward), _p(p), p.shape[0], _p(out))
    return out


def mmd_se_batch(a, b, bw, coords, backward=True, grad=True):
    a, b, c = _f64(a), _f64(b), _f64(np.atleast_2d(coords))
    dim = a.shape[1]
    K, nc = c.shape
    vals = np.zeros(K)
    grads = np.zeros((K, nc)) if grad else None
    rc = lib().cb2o_mmd_se_batch(_p(a), a.shape[0], _p(b), b.shape[0], dim, C.c_double(bw), _p(c), K,
                                 int(backward), _p(vals), _p(grads))
    assert rc == 0
    return (vals, grads) if grad else vals


def kde_eval(mu, sigma, q, w=None, logp=False):
    mu, sigma, q = _f64(mu), _f64(sigma), _f64(q)
    w = None if w is None else _f64(w)
    out = np.zeros(q.shape[0])
    rc = lib().cb2o_kde_eval(_p(mu), _p(sigma), _p(w), mu.shape[1], mu.shape[0], _p(q), q.shape[0], int(logp), _p(out))
    assert rc == 0
    return out


def loo_nll_1d(x, h, circular=False):
    x = _f64(x)
    return lib().cb2o_loo_nll_1d(_p(x), x.size, 1, h, int(circular))


def kde_bw_loo(mu, circular=None):
    mu = _f64(mu)
    N, d = mu.shape
    c = _circ(circular, d)
    out = np.zeros(d)
    rc = lib().cb2o_kde_bw_loo(_p(mu), d, N, _p(c), _p(out))
    assert rc == 0
    return out


def kde_sample(mu, sigma, n, seed, stream=0, w=None, circular=None, ubits=53):
    mu, sigma = _f64(mu), _f64(sigma)
    w = None if w is None else _f64(w)
    N, d = mu.shape
    c = _circ(circular, d)
    out = np.zeros((n, d))
    rc = lib().cb2o_kde_sample(_p(mu), _p(sigma), _p(w), d, N, _p(c), n, C.c_uint64(seed), stream, ubits, _p(out))
    assert rc == 0
    return out


def heatmap_grid_density(img, domain, N, seed):
    """CPU counterpart of `HeatmapGridDensity(img, (x, y); N)` as caesar.jl_b200/factors.py builds it (upstream's
    IIF type is not in the reference tree, R:ext/Images/ScatterAlignPose2.jl:23-24 only calls it): grid cells weighted by the clipped
    image value (cells below 1e-9 of the maximum dropped), N draws from the weighted cell mixture with half a grid spacing of
    Gaussian jitter, LOO bandwidth of the draws.  Returns (points N x 2, bandwidth 2)."""
    img = np.asarray(img, dtype=np.float64)
    x, y = (np.asarray(v, dtype=np.float64) for v in domain)
    cells, w = [], []
    top = max(img.max(), 0.0)
    for i in range(len(x)):
        for j in range(len(y)):
            v = max(img[i, j], 0.0)
            if v > top * 1e-9:
                cells.append((x[i], y[j]))
                w.append(v)
    half = 0.5 * np.array([np.diff(x).mean(), np.diff(y).mean()])
    pts = kde_sample(np.array(cells), half, int(N), seed, w=np.array(w))
    return pts, kde_bw_loo(pts)


def set_level_down(mode):
    """1: a chain descends to a child drawn in proportion to the children's weights (default); 0: always the first child."""
    lib().cb2o_set_level_down(int(mode))


def product(densities, Nout, seed, circular=None, niter=1, stream=0, ubits=53, sample_offset=0):
    """densities: list of (pts N x d, bw d[, w N]).  Returns (pts Nout x d, labels Nout x D)."""
    D = len(densities)
    keep = []
    arr = (Density * D)()
    d = None
    for j, dn in enumerate(densities):
        pts, bw = _f64(dn[0]), _f64(dn[1])
        w = _f64(dn[2]) if len(dn) > 2 and dn[2] is not None else None
        d = pts.shape[1]
        keep += [pts, bw, w]
        arr[j].N = pts.shape[0]
        arr[j].pts = pts.ctypes.data
        arr[j].bw = bw.ctypes.data
        arr[j].w = w.ctypes.data if w is not None else None
    c = _circ(circular, d)
    out = np.zeros((Nout, d))
    labels = np.zeros((Nout, D), dtype=np.int32)
    rc = lib().cb2o_product(arr, D, _p(c), d, Nout, niter, C.c_uint64(seed), stream, ubits, sample_offset,
                            _p(out), _p(labels))
    assert rc == 0
    return out, labels


class Tree:
    """Level hierarchy of one density as the oracle builds it (for device-tree parity tests)."""

    def __init__(self, pts, bw, w=None):
        pts, bw = _f64(pts), _f64(bw)
        w = None if w is None else _f64(w)
        self.N, self.d = pts.shape
        L = lib()
        h = L.cb2o_tree_build(_p(pts), _p(bw), _p(w), self.d, self.N)
        h = C.c_void_p(h)
        self.depth = L.cb2o_tree_depth(h)
        self.perm = np.ctypeslib.as_array(C.cast(L.cb2o_tree_perm(h), C.POINTER(C.c_int32)), (self.N,)).copy()
        self.cnt, self.mean, self.var, self.wt = [], [], [], []
        for l in range(self.depth + 1):
            n = L.cb2o_tree_level_count(h, l)
            self.cnt.append(n)
            dp = C.POINTER(C.c_double)
            self.mean.append(np.ctypeslib.as_array(C.cast(L.cb2o_tree_level_mean(h, l), dp), (n, self.d)).copy())
            self.var.append(np.ctypeslib.as_array(C.cast(L.cb2o_tree_level_var(h, l), dp), (n, self.d)).copy())
            self.wt.append(np.ctypeslib.as_array(C.cast(L.cb2o_tree_level_wt(h, l), dp), (n,)).copy())
        L.cb2o_tree_free(h)


def se_residual(dim, X, p, q):
    """Residual of the relative-pose factors for K particles: vee(log(q, p o exp(X))), product exp / log (coordinates)."""
    X, p, q = (_f64(np.atleast_2d(v)) for v in (X, p, q))
    out = np.zeros_like(X)
    rc = lib().cb2o_se_residual(int(dim), _p(X), _p(p), _p(q), X.shape[0], _p(out))
    assert rc == 0
    return out


def approx_conv(kind, reverse, N, mean, chol, seed, stream=0, src=None, aux=None):
    """Closed-form approxConv (kind 0 Pose2 prior, 1 Pose2Pose2, 2 Pose2Point2BearingRange, 3 Pose3 prior, 4 Pose3Pose3, 5 Point2 prior,
    6 Point2Point2Range, 7 Prior / LinearRelative with caller-sampled noise in `aux`);
    returns N x d_tgt."""
    n = 6 if kind in (3, 4) else 3
    mean, chol = _f64(np.resize(np.asarray(mean, dtype=np.float64), n)), _f64(np.asarray(chol, dtype=np.float64).reshape(n * n))
    src = None if src is None else _f64(src)
    aux = None if aux is None else _f64(aux)
    dt = 6 if kind in (3, 4) else (2 if (kind in (5, 6) or (kind == 2 and not reverse)) else 3)
    if kind == 7:   # linear relative / prior with the caller's noise samples (aux): d = mean[0]
        dt = int(mean[0])
        aux = _f64(np.reshape(aux, (-1, dt)))
        src = None if src is None else _f64(np.reshape(src, (-1, dt)))
    out = np.zeros((N, dt))
    rc = lib().cb2o_approx_conv(kind, int(reverse), N, _p(src), 0 if src is None else src.shape[0], _p(aux),
                                0 if aux is None else aux.shape[0], stream, _p(mean), _p(chol), C.c_uint64(seed), _p(out))
    assert rc == 0
    return out


def mmd_manifold(a, b, circular, bw, rot_weight=2.0):
    a, b = _f64(a), _f64(b)
    c = _circ(circular, a.shape[1])
    out = C.c_double()
    rc = lib().cb2o_mmd_manifold(_p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], _p(c), C.c_double(rot_weight), C.c_double(bw),
                                 C.byref(out))
    assert rc == 0
    return out.value


def icp_align(fix, mov, perturb=None, correspondences=500, neighbors=50, min_planarity=0.3, min_change=3.0, max_iterations=25):
    """simpleICP (R:src/3rdParty/_PCL/services/ICP_Simple.jl:239-332) for K perturbed copies of `mov`; returns (H [K,4,4], stats [K,4])."""
    fix, mov = _f64(fix), _f64(mov)
    K = 1 if perturb is None else len(perturb)
    pert = None if perturb is None else _f64(perturb)
    H, st = np.zeros((K, 4, 4)), np.zeros((K, 4))
    rc = lib().cb2o_icp_align(_p(fix), fix.shape[0], _p(mov), mov.shape[0], _p(pert), K, int(correspondences), int(neighbors),
                              C.c_double(min_planarity), C.c_double(min_change), int(max_iterations), _p(H), _p(st))
    if rc != 0:
        raise ValueError("icp_align: invalid arguments")
    return H, st


def icp_normals(fix, correspondences=500, neighbors=50):
    fix = _f64(fix)
    n = min(correspondences, fix.shape[0])
    sel, nrm, pl = np.zeros(n, dtype=np.int32), np.zeros((n, 3)), np.zeros(n)
    ns = lib().cb2o_icp_normals(_p(fix), fix.shape[0], int(correspondences), int(neighbors), _p(sel), _p(nrm), _p(pl))
    assert ns == n
    return sel, nrm, pl
