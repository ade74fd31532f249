"""Does the LOO bandwidth stage (XU-bound) overlap usefully with the Gibbs kernel (FMA / shared-memory / XU mix) when both run at
once?  Two library contexts on the same GPU, one thread each: A runs the 256-variable Gibbs products without the bandwidth stage,
B the bandwidth searches of a previous result set.  Compares K rounds of both run back to back with K rounds run concurrently."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from __graft_entry__ import load_package
    cb = load_package()
    L = cb._capi.lib()
    V, D, N, d, NO = int(sys.argv[1]) if len(sys.argv) > 1 else 256, bench.D_MSG, bench.N_KER, bench.DIM, bench.N_OUT
    ctxA, ctxB = cb.Context(0, "fp32"), cb.Context(0, "fp32")
    variables = [bench.make_variable(v) for v in range(V)]
    dev_pts = [[torch.from_numpy(p).cuda() for p, _ in var] for var in variables]
    dev_bw = [[torch.from_numpy(b).cuda() for _, b in var] for var in variables]
    dev_out = [torch.empty((NO, d), dtype=torch.float64, device="cuda") for _ in range(V)]
    dev_lab = [torch.empty((NO, D), dtype=torch.int32, device="cuda") for _ in range(V)]
    descs = (cb.ProductDesc * V)()
    keep = []
    for v in range(V):
        arr = (cb.Density * D)(*[cb.Density(N, dev_pts[v][j].data_ptr(), dev_bw[v][j].data_ptr(), None) for j in range(D)])
        keep.append(arr)
        descs[v].dens, descs[v].D, descs[v].Nout, descs[v].niter = arr, D, NO, 1
        descs[v].stream, descs[v].sample_offset = v, 0
        descs[v].pts_out, descs[v].labels_out, descs[v].bw_out = dev_out[v].data_ptr(), dev_lab[v].data_ptr(), None
    circ = np.array(bench.CIRC, dtype=np.uint8)
    prev = [torch.empty((NO, d), dtype=torch.float64, device="cuda") for _ in range(V)]
    sig = torch.empty((V, d), dtype=torch.float64, device="cuda")
    mus = (C.c_void_p * V)(*[p.data_ptr() for p in prev])
    Ns = np.full(V, NO, dtype=np.int32)

    def gibbs(k):
        ctxA.check(L.cb2_product_batch_dev(ctxA._h, descs, V, circ.ctypes.data_as(C.c_void_p), d, 100 + k))
        ctxA.synchronize()

    def loo():
        ctxB.check(L.cb2_kde_bw_loo_batch_dev(ctxB._h, mus, Ns.ctypes.data_as(C.c_void_p), V, d, circ.ctypes.data_as(C.c_void_p), C.c_void_p(sig.data_ptr())))
        ctxB.synchronize()
    gibbs(0)
    for v in range(V):
        prev[v].copy_(dev_out[v])
    torch.cuda.synchronize()
    loo()
    K = 3
    t0 = time.perf_counter()
    for k in range(K):
        gibbs(k); loo()
    seq = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    ta = threading.Thread(target=lambda: [gibbs(k) for k in range(K)])
    tb = threading.Thread(target=lambda: [loo() for _ in range(K)])
    ta.start(); tb.start(); ta.join(); tb.join()
    con = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    for k in range(K):
        gibbs(k)
    g = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    for k in range(K):
        loo()
    l = (time.perf_counter() - t0) / K
    print(f"vars {V}: gibbs alone {g*1e3:.1f} ms, loo alone {l*1e3:.1f} ms, back to back {seq*1e3:.1f} ms, concurrent {con*1e3:.1f} ms per round")


if __name__ == "__main__":
    main()
