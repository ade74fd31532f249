"""Per-call cost of the device-resident entry points on solver-sized problems (3 Pose2 products of 2 messages, N = 100 / 1000):
100 back-to-back calls, one synchronisation at the end -> max(host time, device time) per call."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

import torch
cb = load_package()
ctx = cb.Context(0, "fp32")
L = cb._capi.lib()
rng = np.random.default_rng(0)
for N in (100, 1000):
    B, D, d = 3, 2, 3
    pts = [[torch.from_numpy(rng.normal(size=(N, d))).cuda() for _ in range(D)] for _ in range(B)]
    bws = [[torch.from_numpy(np.full(d, 0.3)).cuda() for _ in range(D)] for _ in range(B)]
    out = [torch.empty((N, d), dtype=torch.float64, device="cuda") for _ in range(B)]
    bwo = [torch.empty(d, dtype=torch.float64, device="cuda") for _ in range(B)]
    descs = (cb.ProductDesc * B)()
    keep = []
    for b in range(B):
        arr = (cb.Density * D)(*[cb.Density(N, pts[b][j].data_ptr(), bws[b][j].data_ptr(), None) for j in range(D)])
        keep.append(arr)
        descs[b].dens, descs[b].D, descs[b].Nout, descs[b].niter, descs[b].stream = arr, D, N, 1, b
        descs[b].pts_out, descs[b].labels_out, descs[b].bw_out = out[b].data_ptr(), None, bwo[b].data_ptr()
    circ = np.array([0, 0, 1], dtype=np.uint8)
    cp = circ.ctypes.data_as(C.c_void_p)
    for with_bw in (True, False):
        for b in range(B):
            descs[b].bw_out = bwo[b].data_ptr() if with_bw else None
        for _ in range(5):
            ctx.check(L.cb2_product_batch_dev(ctx._h, descs, B, cp, d, 1))
        ctx.synchronize()
        t0 = time.perf_counter()
        for k in range(100):
            ctx.check(L.cb2_product_batch_dev(ctx._h, descs, B, cp, d, k))
        t1 = time.perf_counter()
        ctx.synchronize()
        t2 = time.perf_counter()
        print(f"N={N} product_batch_dev x3 (bw_out={with_bw}): host enqueue {1e4*(t1-t0):.1f} us/call, with final sync {1e4*(t2-t0):.1f} us/call")
    mus = (C.c_void_p * B)(*[o.data_ptr() for o in out])
    Ns = np.full(B, N, dtype=np.int32)
    sig = torch.empty((B, d), dtype=torch.float64, device="cuda")
    for _ in range(5):
        ctx.check(L.cb2_kde_bw_loo_batch_dev(ctx._h, mus, Ns.ctypes.data_as(C.c_void_p), B, d, cp, C.c_void_p(sig.data_ptr())))
    ctx.synchronize()
    t0 = time.perf_counter()
    for k in range(100):
        ctx.check(L.cb2_kde_bw_loo_batch_dev(ctx._h, mus, Ns.ctypes.data_as(C.c_void_p), B, d, cp, C.c_void_p(sig.data_ptr())))
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print(f"N={N} kde_bw_loo_batch_dev x3: host enqueue {1e4*(t1-t0):.1f} us/call, with final sync {1e4*(t2-t0):.1f} us/call")
