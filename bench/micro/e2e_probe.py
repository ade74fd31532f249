"""Where the host-pointer product call spends its time beyond the kernels (run on the GPU box): stage timers of one C5 call,
and the plain pinned H2D / D2H bandwidth of the same buffers through torch for comparison."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

cb = load_package()
V = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = cb.Context(0, "fp32")
L = cb._capi.lib()
variables = [bench.make_variable(v) for v in range(V)]
pin_pts = [[torch.from_numpy(p).pin_memory() for p, _ in var] for var in variables]
pin_out = [torch.empty((bench.N_OUT, bench.DIM), dtype=torch.float64).pin_memory() for _ in range(V)]
pin_lab = [torch.empty((bench.N_OUT, bench.D_MSG), dtype=torch.int32).pin_memory() for _ in range(V)]
pin_bw = [torch.zeros(bench.DIM, dtype=torch.float64).pin_memory() for _ in range(V)]
keep = []
hd = (cb.ProductDesc * V)()
for v in range(V):
    arr = (cb.Density * bench.D_MSG)(*[cb.Density(bench.N_KER, pin_pts[v][j].data_ptr(), variables[v][j][1].ctypes.data, None) for j in range(bench.D_MSG)])
    keep.append(arr)
    hd[v].dens, hd[v].D, hd[v].Nout, hd[v].niter, hd[v].stream = arr, bench.D_MSG, bench.N_OUT, 1, v
    hd[v].pts_out, hd[v].labels_out, hd[v].bw_out = pin_out[v].data_ptr(), pin_lab[v].data_ptr(), pin_bw[v].data_ptr()
circ = np.array(bench.CIRC, dtype=np.uint8)
for it in range(4):
    t0 = time.perf_counter()
    ctx.check(L.cb2_product_batch(ctx._h, hd, V, circ.ctypes.data_as(C.c_void_p), bench.DIM, 7 + it))
    dt = time.perf_counter() - t0
    st = {k: L.cb2_last_stage_ms(ctx._h, k.encode()) for k in ("h2d", "tree", "gibbs", "bw", "d2h")}
    print(f"call {it}: {dt * 1e3:.1f} ms host;", {k: round(v, 2) for k, v in st.items()}, "sum", round(sum(st.values()), 1))
dev = [torch.empty_like(t, device="cuda") for var in pin_pts for t in var]
torch.cuda.synchronize()
t0 = time.perf_counter()
for d_, s_ in zip(dev, [t for var in pin_pts for t in var]):
    d_.copy_(s_, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
nb = sum(t.numel() * 8 for var in pin_pts for t in var)
print(f"torch pinned H2D of the same {len(dev)} arrays: {dt * 1e3:.1f} ms = {nb / dt / 1e9:.1f} GB/s")
big = torch.empty(nb // 8, dtype=torch.float64).pin_memory()
dbig = torch.empty_like(big, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"one {nb / 1e6:.0f} MB pinned H2D copy: {dt * 1e3:.1f} ms = {nb / dt / 1e9:.1f} GB/s")
