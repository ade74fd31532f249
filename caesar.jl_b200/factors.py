"""ScatterAlignPose2 / ScatterAlignPose3 factor plugin, mirroring the reference's IIF factor interface
(struct R:ext/factors/ScatterAlignPose.jl:50-97; preambleCache R:ext/Images/ScatterAlignPose2.jl:115-152;
getSample :155-187).  The alignment itself (fresh samples from both clouds, mmd cost, BFGS) runs in
libcaesar_b200.so via cb2_scatter_align; K particles of the solver are aligned in one batched call."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import f64, ptr


@dataclass
class ScatterAlign:
    cloud1: object
    cloud2: object
    dim: int = 2
    gridscale: float = 1.0
    sample_count: int = 500
    bw: float = 1.0
    useStashing: bool = False
    dataEntry_cloud1: str = ""
    dataEntry_cloud2: str = ""
    dataStoreHint: str = ""


class _ScatterAlignPose:
    DIM = 2

    def __init__(self, cloud1=None, cloud2=None, sample_count=75, bw=5e-5, rescale=1.0, useStashing=False,
                 dataEntry_cloud1="", dataEntry_cloud2="", dataStoreHint=""):
        # keyword defaults as the reference constructors (R:ext/Images/ScatterAlignPose2.jl:36-46)
        if cloud1 is None or cloud2 is None:
            raise ValueError("cloud1 and cloud2 are required")
        if cloud1.d != self.DIM or cloud2.d != self.DIM:
            raise ValueError(f"{type(self).__name__} needs clouds over R^{self.DIM}")
        self.align = ScatterAlign(cloud1, cloud2, self.DIM, float(rescale), int(sample_count), float(bw), useStashing,
                                  str(dataEntry_cloud1), str(dataEntry_cloud2), str(dataStoreHint))

    @property
    def ncoord(self):
        return 3 if self.DIM == 2 else 6


class ScatterAlignPose2(_ScatterAlignPose):
    DIM = 2


class ScatterAlignPose3(_ScatterAlignPose):
    DIM = 3


@dataclass
class CalcFactor:
    """The slice of IIF.CalcFactor the factor uses: .factor, .cache, ._allowThreads."""
    factor: object
    cache: dict = field(default_factory=dict)
    _allowThreads: bool = True


def preambleCache(dfg, vars, fnc):
    """Runs once at addFactor!.  Mirrors the named tuple (;M,e0,smps1,smps2,bw,score) of the reference; the
    sample buffers live on the device inside cb2_scatter_align, so only sizes are recorded here."""
    # sample_count < 0 selects the ICP branch: no sample buffers, the clouds are aligned as they are
    M = "SpecialEuclidean(2)" if fnc.DIM == 2 else "SpecialEuclidean(3)"
    return dict(M=M, e0=np.zeros(fnc.ncoord), smps1=None, smps2=None, bw=np.array([fnc.align.bw]), score=[0.0])


def getSample(cf, K=1, seed=0, x0=None, max_iter=200, g_tol=1e-8, ctx=None, return_info=False):
    """getSample(cf::CalcFactor{ScatterAlignPose2/3}): draws sample_count fresh points from each cloud and
    minimises mmd(smps1, T(c)^-1 smps2) over the SE(2)/SE(3) coordinates c, from the reference's random start
    [5 randn(ntr); 0.1 randn(nrot)] (:182).  Returns the K x ncoord tangent coordinates (hat(M, e0, c) in the
    reference is a re-labelling of the same numbers) and stores the last score in cf.cache['score']."""
    from . import context
    ctx = ctx or context()
    fnc = cf.factor
    al = fnc.align
    nc, dim = fnc.ncoord, fnc.DIM
    if al.sample_count < 0:
        return _getSampleICP(cf, K, seed, ctx, return_info)
    if x0 is None:
        rng = np.random.default_rng(seed)
        x0 = np.concatenate([5.0 * rng.standard_normal((K, dim)), 0.1 * rng.standard_normal((K, nc - dim))], axis=1)
    x0 = f64(np.atleast_2d(x0))
    K = x0.shape[0]
    coords, score = np.zeros((K, nc)), np.zeros(K)
    iters = np.zeros(K, dtype=np.int32)
    d1, d2 = al.cloud1._density(), al.cloud2._density()
    ctx.check(_capi.lib().cb2_scatter_align(ctx._h, C.byref(d1), C.byref(d2), dim, al.sample_count, al.bw, ptr(x0), K,
                                            int(seed), int(max_iter), float(g_tol), ptr(coords), ptr(score), ptr(iters)))
    cf.cache.setdefault("score", [0.0])[0] = float(score[-1])
    if return_info:
        return coords, score, iters
    return coords


def _se3_log_coords(H):
    """log(M, e0, ArrayPartition(t, R)) of SpecialEuclidean(3) as tangent coordinates [t; rotation vector] (:211-213)."""
    R, t = H[:3, :3], H[:3, 3]
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    w = w * (th / np.sin(th)) if th > 1e-9 else w
    return np.r_[t, w]


def _getSampleICP(cf, K, seed, ctx, return_info):
    """sample_count < 0 (R:ext/Images/ScatterAlignPose2.jl:187-214): every particle perturbs cloud2 by dstb = 0.2 randn(6)
    (backward) and aligns it to cloud1 with simpleICP (max_iterations = 25, correspondences = 500, neighbors = 50); the K
    particles run as one device batch.  Returns the K x 6 tangent coordinates of fix_H_mov."""
    from . import icp_align_batch
    fnc = cf.factor
    if fnc.DIM != 3:
        raise NotImplementedError("ICP alignments are implemented for Pose3 (three-dimensional clouds), as upstream")
    al = fnc.align
    rng = np.random.default_rng(seed)
    dstb = 0.2 * rng.standard_normal((K, 6))
    H, st = icp_align_batch(al.cloud1.pts, al.cloud2.pts, dstb, correspondences=500, neighbors=50, max_iterations=25, ctx=ctx)
    coords = np.array([_se3_log_coords(h) for h in H])
    cf.cache.setdefault("score", [0.0])[0] = float(st[-1, 0])
    if return_info:
        return coords, st[:, 0], st[:, 3].astype(np.int32)
    return coords


class HeatmapGridDensity:
    """IIF.HeatmapGridDensity(img, (x, y); N) as ScatterAlignPose2 uses it (R:ext/Images/ScatterAlignPose2.jl:23-24,
    R:test/testScatterAlignPose2.jl:22-53): a grid image turned into a sampled belief.  Construction only (SURVEY
    sec. 8a row a10): N points are drawn from the grid cells in proportion to the image values (jittered by the grid
    spacing) with the device sampler, then wrapped as a ManifoldKernelDensity (`densityFnc`) with a LOO bandwidth."""

    def __init__(self, img, domain, N=1000, seed=0, ctx=None):
        from . import ManifoldKernelDensity, manikde, sample
        img = np.asarray(img, dtype=np.float64)
        x, y = (np.asarray(v, dtype=np.float64) for v in domain)
        if img.shape != (len(x), len(y)):
            raise ValueError("image shape does not match the domain")
        w = np.clip(img, 0.0, None).ravel()
        if not w.sum() > 0:
            raise ValueError("heatmap has no positive mass")
        keep = w > w.max() * 1e-9
        gx, gy = np.meshgrid(x, y, indexing="ij")
        cells = np.c_[gx.ravel()[keep], gy.ravel()[keep]]
        spacing = np.array([np.mean(np.diff(x)), np.mean(np.diff(y))])
        grid = ManifoldKernelDensity("Point2", cells, spacing / 2.0, weights=w[keep])
        pts, = sample(grid, int(N), seed=seed, ctx=ctx)
        self.data, self.domain, self.N = img, (x, y), int(N)
        self.densityFnc = manikde("Point2", pts, ctx=ctx)

    # the factor treats a heatmap cloud exactly like the MKD it wraps
    def __getattr__(self, name):
        return getattr(self.__dict__["densityFnc"], name)


def ScatterAlignPose2_from_images(im1, im2, domain, sample_count=75, bw=5e-5, N=1000, seed=0, ctx=None, **kw):
    """ScatterAlignPose2(im1, im2, domain; sample_count, bw, N) -- the image constructor of the reference
    (R:ext/Images/ScatterAlignPose2.jl:4-34; `cvt`/`rescale` image resizing is left to the caller)."""
    c1 = HeatmapGridDensity(im1, domain, N=N, seed=seed, ctx=ctx)
    c2 = HeatmapGridDensity(im2, domain, N=N, seed=seed + 1, ctx=ctx)
    return ScatterAlignPose2(cloud1=c1.densityFnc, cloud2=c2.densityFnc, sample_count=sample_count, bw=bw, **kw)
