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
    """The slice of IIF.CalcFactor the factor uses: .factor, .cache, ._allowThreads.  Calling it evaluates the factor
    residual `(cf::CalcFactor{<:ScatterAlignPose2/3})(X, p, q)` (R:ext/Images/ScatterAlignPose2.jl:216-231)."""
    factor: object
    cache: dict = field(default_factory=dict)
    _allowThreads: bool = True

    def __call__(self, X, p, q, ctx=None):
        """X: tangent coordinates of the measurement at the identity, p / q: pose coordinates ([x, y, theta] or
        [t; rotation vector]); one particle (1-D arrays) or K particles (K x ncoord).  Returns
        `vee(M, q, log(M, q, p o exp(M, e0, X)))` -- computed for all particles in one device launch (cb2_se_residual_batch)."""
        from . import se_residual_batch
        X = np.asarray(X, dtype=np.float64)
        one = X.ndim == 1
        out = se_residual_batch(self.factor.DIM, np.atleast_2d(X), np.atleast_2d(p), np.atleast_2d(q), ctx=ctx)
        return out[0] if one else out


def getManifold(fnc):
    """getManifold(::ScatterAlignPose2) = getManifold(Pose2Pose2), Pose3 likewise (R:ext/Images/ScatterAlignPose2.jl:111-112)."""
    return "SpecialEuclidean(2)" if fnc.DIM == 2 else "SpecialEuclidean(3)"


def preambleCache(dfg, vars, fnc):
    """Runs once at addFactor! (R:ext/Images/ScatterAlignPose2.jl:115-152).  With `useStashing` the two clouds are first
    reconstituted from the blob store: `getData(dfg, label, UUID(dataEntry))` returns the JSON of a PackedManifoldKernelDensity,
    which replaces the (parched) cloud the factor was unpacked with (:124-137).  Mirrors the named tuple
    (;M,e0,smps1,smps2,bw,score) of the reference; the sample buffers live on the device inside cb2_scatter_align, so only
    sizes are recorded here."""
    al = fnc.align
    if al.useStashing:
        from .serialization import unpackDistribution
        if dfg is None or not hasattr(dfg, "getData"):
            raise ValueError("useStashing needs a graph with a blob store (getData)")
        for va, de, which in zip(vars, (al.dataEntry_cloud1, al.dataEntry_cloud2), ("cloud1", "cloud2")):
            if not de:
                raise AssertionError(f"cannot reconstitute {type(fnc).__name__} without necessary data entry, only have {de!r}")
            _, blob = dfg.getData(getattr(va, "label", va), de)
            cld = unpackDistribution(blob.decode() if isinstance(blob, (bytes, bytearray)) else blob)
            _update(getattr(al, which), cld)
    # sample_count < 0 selects the ICP branch: no sample buffers, the clouds are aligned as they are
    return dict(M=getManifold(fnc), e0=np.zeros(fnc.ncoord), smps1=None, smps2=None, bw=np.array([al.bw]), score=[0.0])


def _update(dst, src):
    """`_update!(cl, cld)`: the stored belief takes the points and bandwidth of the reconstituted one, in place."""
    dst.pts, dst.bw = src.pts, src.bw
    dst.N, dst.d = src.pts.shape
    dst.weights = src.weights


def parchDistribution(mkd):
    """IIF.parchDistribution(::ManifoldKernelDensity): the belief reduced to its first point (same bandwidth, partial,
    infoPerCoord) -- what a stashed factor carries instead of the full cloud (R:ext/Images/ScatterAlignPose2.jl:362-367)."""
    from . import ManifoldKernelDensity
    return ManifoldKernelDensity(mkd.manifold, mkd.pts[:1].copy(), mkd.bw.copy(), None, mkd.partial, mkd.infoPerCoord)


_PACKED_TYPE = {2: "Caesar.PackedScatterAlignPose2", 3: "Caesar.PackedScatterAlignPose3"}


def packScatterAlign(arp):
    """`convert(PackedScatterAlignPose2/3, arp)` (R:ext/Images/ScatterAlignPose2.jl:354-385; fields
    R:ext/factors/ScatterAlignPose.jl:100-132): both clouds as PackedManifoldKernelDensity dictionaries -- parched to a single
    point when `useStashing`, in which case both data-entry identifiers must be present."""
    from .serialization import packDistribution
    al = arp.align
    cld1, cld2 = al.cloud1, al.cloud2
    if al.useStashing:
        if not al.dataEntry_cloud1:
            raise AssertionError("packing of ScatterAlignPose asked to be `useStashing=true` yet no `.dataEntry_cloud1` exists for later loading")
        if not al.dataEntry_cloud2:
            raise AssertionError("packing of ScatterAlignPose asked to be `useStashing=true` yet no `.dataEntry_cloud2` exists for later loading")
        cld1, cld2 = parchDistribution(cld1), parchDistribution(cld2)
    return {"_type": _PACKED_TYPE[arp.DIM], "cloud1": packDistribution(cld1), "cloud2": packDistribution(cld2),
            "gridscale": al.gridscale, "sample_count": al.sample_count, "bw": al.bw, "useStashing": bool(al.useStashing),
            "dataEntry_cloud1": al.dataEntry_cloud1, "dataEntry_cloud2": al.dataEntry_cloud2, "dataStoreHint": al.dataStoreHint}


def unpackScatterAlign(parp):
    """`convert(ScatterAlignPose2/3, parp)` (R:ext/Images/ScatterAlignPose2.jl:387-427).  Accepts the dictionary or its JSON."""
    import json
    from .serialization import unpackDistribution
    if isinstance(parp, (str, bytes)):
        parp = json.loads(parp)
    kinds = {v: k for k, v in _PACKED_TYPE.items()}
    if parp.get("_type") not in kinds:
        raise ValueError(f"unsupported packed factor type {parp.get('_type')!r}")
    cls = ScatterAlignPose2 if kinds[parp["_type"]] == 2 else ScatterAlignPose3
    return cls(cloud1=unpackDistribution(parp["cloud1"]), cloud2=unpackDistribution(parp["cloud2"]),
               sample_count=parp.get("sample_count", 50), bw=parp.get("bw", 0.01), rescale=parp.get("gridscale", 1.0),
               useStashing=parp.get("useStashing", False), dataEntry_cloud1=parp.get("dataEntry_cloud1", ""),
               dataEntry_cloud2=parp.get("dataEntry_cloud2", ""), dataStoreHint=parp.get("dataStoreHint", ""))


class BlobStoreGraph:
    """The two DFG calls the stashed factor needs (R:test/testScatterAlignParched.jl:62-80,:118-127): `addData!` stores a blob
    under a variable label and returns its entry id, `getData` fetches it back.  In-memory stand-in for a FolderStore."""

    def __init__(self):
        self._blobs = {}

    def addData(self, label, key, blob):
        import uuid
        eid = str(uuid.uuid5(uuid.NAMESPACE_OID, f"{label}/{key}/{len(self._blobs)}"))
        self._blobs[(str(label), eid)] = (key, bytes(blob))
        return eid

    def getData(self, label, entry_id):
        key, blob = self._blobs[(str(label), str(entry_id))]
        return key, blob


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
