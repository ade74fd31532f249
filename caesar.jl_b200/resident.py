"""Device-resident beliefs: the host-side mirror of the `*_dev` entry points.

A belief's points (N x d Float64, point-contiguous as in Julia's d x N) and bandwidths (d) live in device memory between
approxConvBelief -> manikde! -> manifoldProduct -> manikde! (IIF `propagateBelief`); the calls are stream-ordered and
never wait for the host.  `DeviceBelief.to_host()` is the only synchronisation (what `getPoints` /
`packDistribution` would trigger on the Julia side).  No numerics happen here: every operation is one C-ABI call.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import ConvDesc, Density, ProductDesc, circ_arr, f64, ptr


class DeviceSlab:
    """One device allocation carved up with a bump pointer (cudaMalloc per belief would serialise the stream)."""

    def __init__(self, ctx, nbytes=64 << 20):
        self.ctx, self.cap, self.off = ctx, int(nbytes), 0
        p = C.c_void_p()
        ctx.check(_capi.lib().cb2_device_alloc(ctx._h, self.cap, C.byref(p)))
        self.base = p.value

    def take(self, nbytes):
        o = (self.off + 255) // 256 * 256
        if o + nbytes > self.cap:
            raise MemoryError(f"device slab of {self.cap} bytes exhausted")
        self.off = o + int(nbytes)
        return self.base + o

    def reset(self):
        self.off = 0

    def free(self):
        if self.base:
            self.ctx.check(_capi.lib().cb2_device_free(self.ctx._h, C.c_void_p(self.base)))
            self.base = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceArray:
    """N x d Float64 samples on the device (a proposal before it has a bandwidth)."""
    __slots__ = ("ptr", "N", "d")

    def __init__(self, ptr_, N, d):
        self.ptr, self.N, self.d = int(ptr_), int(N), int(d)


class DeviceBelief:
    """ManifoldKernelDensity whose points and bandwidths stay on the device."""
    __slots__ = ("manifold", "circular", "N", "d", "pts", "bw_ptr", "ctx")

    def __init__(self, manifold, circular, pts, bw_ptr, ctx):
        self.manifold, self.circular, self.pts, self.bw_ptr, self.ctx = manifold, tuple(circular), pts, int(bw_ptr), ctx
        self.N, self.d = pts.N, pts.d

    def _density(self):
        return Density(self.N, self.pts.ptr, self.bw_ptr, None)

    def to_host(self):
        """Downloads the belief (waits for the stream) -> ManifoldKernelDensity."""
        from . import ManifoldKernelDensity
        L = _capi.lib()
        pts, bw = np.zeros((self.N, self.d)), np.zeros(self.d)
        self.ctx.check(L.cb2_download(self.ctx._h, ptr(pts), C.c_void_p(self.pts.ptr), pts.nbytes))
        self.ctx.check(L.cb2_download(self.ctx._h, ptr(bw), C.c_void_p(self.bw_ptr), bw.nbytes))
        return ManifoldKernelDensity(self.manifold, pts, bw)


def download_beliefs(beliefs, ctx):
    """DeviceBelief list -> ManifoldKernelDensity list with ONE device-to-host copy (the address range that holds them all; the
    beliefs of a finished solve sit next to each other in the slab) instead of two synchronous copies per belief."""
    from . import ManifoldKernelDensity
    if not beliefs:
        return []
    spans = [(b.pts.ptr, b.N * b.d * 8) for b in beliefs] + [(b.bw_ptr, b.d * 8) for b in beliefs]
    lo = min(p for p, _ in spans)
    hi = max(p + n for p, n in spans)
    if hi - lo > (64 << 20):   # scattered over a large range: not worth one big copy
        return [b.to_host() for b in beliefs]
    host = np.empty(hi - lo, dtype=np.uint8)
    ctx.check(_capi.lib().cb2_download(ctx._h, ptr(host), C.c_void_p(lo), host.nbytes))
    out = []
    for b in beliefs:
        pts = host[b.pts.ptr - lo: b.pts.ptr - lo + b.N * b.d * 8].view(np.float64).reshape(b.N, b.d).copy()
        bw = host[b.bw_ptr - lo: b.bw_ptr - lo + b.d * 8].view(np.float64).copy()
        out.append(ManifoldKernelDensity(b.manifold, pts, bw))
    return out


def upload_array(arr, slab, ctx):
    """Host N x d Float64 samples -> DeviceArray (one stream-ordered copy)."""
    a = f64(np.atleast_2d(arr))
    p = slab.take(a.nbytes)
    ctx.check(_capi.lib().cb2_upload(ctx._h, C.c_void_p(p), ptr(a), a.nbytes))
    return DeviceArray(p, a.shape[0], a.shape[1])


def upload_belief(mkd, slab, ctx):
    """Host ManifoldKernelDensity -> DeviceBelief (two stream-ordered copies)."""
    L = _capi.lib()
    pts, bw = f64(mkd.pts), f64(mkd.bw)
    pp, bp = slab.take(pts.nbytes), slab.take(bw.nbytes)
    ctx.check(L.cb2_upload(ctx._h, C.c_void_p(pp), ptr(pts), pts.nbytes))
    ctx.check(L.cb2_upload(ctx._h, C.c_void_p(bp), ptr(bw), bw.nbytes))
    return DeviceBelief(mkd.manifold, mkd.circular, DeviceArray(pp, mkd.N, mkd.d), bp, ctx)


def approxConvBatch_dev(convs, slab, seed=0, ctx=None):
    """`approxConvBatch` with device-resident sources: convs[i]['src'] / ['aux'] are DeviceArray (or None); returns DeviceArray."""
    from . import _conv_dims
    B = len(convs)
    descs = (ConvDesc * B)()
    outs = []
    for b, c in enumerate(convs):
        kind, rev = int(c["kind"]), int(bool(c.get("reverse", False)))
        nz, dt = _conv_dims(kind, rev, c.get("mean"))
        N = int(c["N"])
        out = DeviceArray(slab.take(N * dt * 8), N, dt)
        outs.append(out)
        d = descs[b]
        d.kind, d.reverse, d.N, d.stream = kind, rev, N, int(c.get("stream", b))
        src, aux = c.get("src"), c.get("aux")
        d.n_src = 0 if src is None else src.N
        d.n_aux = 0 if aux is None else aux.N
        d.src = None if src is None else src.ptr
        d.aux = None if aux is None else aux.ptr
        # (slice assignment fills a ctypes array in one call: the element loop cost more than the kernel it describes)
        d.mean[:nz] = np.resize(np.asarray(c["mean"], dtype=np.float64), nz).tolist()
        d.chol[:nz * nz] = np.asarray(c["chol"], dtype=np.float64).reshape(nz * nz).tolist()
        d.out = out.ptr
    ctx.check(_capi.lib().cb2_approx_conv_batch_dev(ctx._h, descs, B, int(seed)))
    return outs


def manikde_dev(manifold, arrays, slab, ctx):
    """`manikde!(M, pts)` for a batch of device-resident sample sets: LOO bandwidths computed and kept on the device."""
    from . import _topology
    B = len(arrays)
    d = arrays[0].d
    topo = _topology(manifold, d)
    circ = circ_arr(topo, d)
    mus = (C.c_void_p * B)(*[a.ptr for a in arrays])
    Ns = np.array([a.N for a in arrays], dtype=np.int32)
    bw_base = slab.take(B * d * 8)
    ctx.check(_capi.lib().cb2_kde_bw_loo_batch_dev(ctx._h, mus, ptr(Ns), B, d, ptr(circ), C.c_void_p(bw_base)))
    return [DeviceBelief(manifold, topo, a, bw_base + b * d * 8, ctx) for b, a in enumerate(arrays)]


def manifoldProducts_dev(groups, manifold, N, slab, Niter=1, seed=0, streams=None, ctx=None):
    """Batched `manifoldProduct` over device-resident beliefs; results (points + fresh LOO bandwidth) stay on the device.
    A product of one density is that density (upstream early-out)."""
    from . import _topology
    B = len(groups)
    d = groups[0][0].d
    topo = _topology(manifold, d)
    streams = list(range(B)) if streams is None else streams
    results = [g[0] if len(g) == 1 else None for g in groups]
    multi = [b for b in range(B) if len(groups[b]) > 1]
    if multi:
        descs = (ProductDesc * len(multi))()
        keep = []
        bw_base = slab.take(len(multi) * d * 8)
        for i, b in enumerate(multi):
            g = groups[b]
            arr = (Density * len(g))(*[m._density() for m in g])
            keep.append(arr)
            out = DeviceArray(slab.take(int(N) * d * 8), N, d)
            descs[i].dens, descs[i].D, descs[i].Nout, descs[i].niter = arr, len(g), int(N), int(Niter)
            descs[i].stream, descs[i].sample_offset = int(streams[b]), 0
            descs[i].pts_out, descs[i].labels_out, descs[i].bw_out = out.ptr, None, bw_base + i * d * 8
            results[b] = DeviceBelief(manifold, topo, out, bw_base + i * d * 8, ctx)
        circ = circ_arr(topo, d)
        ctx.check(_capi.lib().cb2_product_batch_dev(ctx._h, descs, len(multi), ptr(circ), d, int(seed)))
    return results
