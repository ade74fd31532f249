"""Times tree_build + gibbs_kernel of one library build on the C5 shape (device-resident inputs, no bandwidth stage) and checks its
chains against the CPU oracle: `CB2_LIB=caesar.jl_b200/variants/lib_X.so python bench/micro/gibbs_variants.py --vars 64`.
One JSON line per run; used to compare experiment builds (bench/micro/build_variant.py) inside one gpurun call."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=256, help="chains of variable 0 compared with the oracle (0: skip)")
    ap.add_argument("--bw", action="store_true", help="also run the LOO bandwidth stage")
    args = ap.parse_args()
    import torch
    import bench
    from __graft_entry__ import load_package
    cb = load_package()
    ctx = cb.Context(0, "fp32")
    L = cb._capi.lib()
    V, D, N, d, NO = args.vars, bench.D_MSG, bench.N_KER, bench.DIM, bench.N_OUT
    variables = [bench.make_variable(v) for v in range(V)]
    dev_pts = [[torch.from_numpy(p).cuda() for p, _ in var] for var in variables]
    dev_bw = [[torch.from_numpy(b).cuda() for _, b in var] for var in variables]
    dev_out = [torch.empty((NO, d), dtype=torch.float64, device="cuda") for _ in range(V)]
    dev_lab = [torch.empty((NO, D), dtype=torch.int32, device="cuda") for _ in range(V)]
    dev_bwo = [torch.empty(d, dtype=torch.float64, device="cuda") for _ in range(V)]
    descs = (cb.ProductDesc * V)()
    keep = []
    for v in range(V):
        arr = (cb.Density * D)(*[cb.Density(N, dev_pts[v][j].data_ptr(), dev_bw[v][j].data_ptr(), None) for j in range(D)])
        keep.append(arr)
        descs[v].dens, descs[v].D, descs[v].Nout, descs[v].niter = arr, D, NO, 1
        descs[v].stream, descs[v].sample_offset = v, 0
        descs[v].pts_out, descs[v].labels_out = dev_out[v].data_ptr(), dev_lab[v].data_ptr()
        descs[v].bw_out = dev_bwo[v].data_ptr() if args.bw else None
    circ = np.array(bench.CIRC, dtype=np.uint8)
    torch.cuda.synchronize()
    best = {}
    for it in range(args.reps + 1):
        ctx.check(L.cb2_product_batch_dev(ctx._h, descs, V, circ.ctypes.data_as(C.c_void_p), d, 100 + it))
        ctx.synchronize()
        if it:
            for k in ("tree", "gibbs", "bw"):
                ms = ctx.stage_ms(k)
                if ms > 0:
                    best[k] = min(best.get(k, 1e30), ms)
    out = {"lib": os.path.basename(cb._capi.SO_PATH), "vars": V, **{k + "_ms": round(v, 3) for k, v in best.items()}}
    out["gibbs_ms_per_256"] = round(best["gibbs"] * 256 / V, 2)
    if args.check:
        import oracle
        oracle.use_all_cores()
        seed = 100 + args.reps
        ref_pts, ref_lab = oracle.product(variables[0], args.check, seed=seed, circular=list(bench.CIRC), ubits=24, stream=0)
        lab = dev_lab[0][:args.check].cpu().numpy()
        pts = dev_out[0][:args.check].cpu().numpy()
        same = (lab == ref_lab).all(axis=1)
        out["label_agreement"] = float(same.mean())
        out["per_message_agreement"] = float((lab == ref_lab).mean())
        if same.any():
            dp = np.abs(pts[same] - ref_pts[same])
            dp[:, 3:] = np.minimum(dp[:, 3:], 2 * np.pi - dp[:, 3:])
            out["max_point_diff_same_labels"] = float(dp.max())
        out["finite"] = bool(torch.isfinite(dev_out[0]).all().item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
