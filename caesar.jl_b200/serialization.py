"""PackedManifoldKernelDensity wire format (the JSON the reference stores for beliefs and stashed clouds:
fixtures R:test/ICRA2022_tests/test_data/tut3/*.json; writer R:test/ICRA2022_tests/insufficient_data.jl:34-44)."""
import json

import numpy as np


def packDistribution(mkd, varType=None):
    if hasattr(mkd, "to_host"):  # device-resident belief (resident.DeviceBelief): packing is where it is read back
        mkd = mkd.to_host()
    vt = varType or (mkd.manifold if isinstance(mkd.manifold, str) else "ContinuousEuclid")
    if "." not in vt:
        vt = "RoME." + vt
    return {
        "_type": "IncrementalInference.PackedManifoldKernelDensity",
        "varType": vt,
        "pts": [list(map(float, p)) for p in mkd.pts],
        "bw": list(map(float, mkd.bw)),
        "partial": list(mkd.partial) if mkd.partial is not None else list(range(1, mkd.d + 1)),
        "infoPerCoord": list(map(float, mkd.infoPerCoord)),
    }


def unpackDistribution(packed):
    from . import ManifoldKernelDensity
    if isinstance(packed, (str, bytes)):
        packed = json.loads(packed)
    if packed.get("_type") != "IncrementalInference.PackedManifoldKernelDensity":
        raise ValueError(f"unsupported packed type {packed.get('_type')!r}")
    vt = packed["varType"].split(".")[-1]
    pts = np.asarray(packed["pts"], dtype=np.float64)
    partial = packed.get("partial")
    if partial is not None and list(partial) == list(range(1, pts.shape[1] + 1)):
        partial = None
    return ManifoldKernelDensity(vt, pts, packed["bw"], None, partial, packed.get("infoPerCoord"))


def multiplyDistributions(requestDict, seed=0, ctx=None):
    """The ZMQ verb `multiplyDistributions` (R:ext/services/additional.jl:11-32) on the device product path:
    every input `Packed_BallTreeDensity` (R:ext/models/distributions.jl:26-32; 1-D, points stacked in a flat list) gets
    a fresh LOO bandwidth (the reference recomputes it too, :18-20), the beliefs are multiplied with
    `manifoldProduct(ContinuousScalar, funcs)` and the reply carries `bw`, equal weights and the new points."""
    from . import manikde, manifoldProduct
    payload = requestDict["payload"]
    funcs = []
    for pbd in payload["weights"]:
        dim = int(pbd.get("dim", 1))
        if dim != 1:
            raise ValueError("multiplyDistributions assumes 1-D Euclidean beliefs, as the reference does")
        pts = np.asarray(pbd["points"], dtype=np.float64).reshape(-1, 1)
        funcs.append(manikde("ContinuousScalar", pts, ctx=ctx))
    pqr = manifoldProduct(funcs, "ContinuousScalar", seed=seed, ctx=ctx)
    n = pqr.N
    resObj = {"dim": 1, "kernelType": "Gaussian", "bandwidth": [float(b) for b in pqr.bw],
              "weights": [1.0 / n] * n, "points": [float(x) for x in pqr.pts[:, 0]]}
    return {"status": "OK", "payload": resObj}


def solveTree(configDict, fg, requestDict, backend=None, gibbs_iters=3, seed=0):
    """The ZMQ verb `solveTree` (R:ext/services/session.jl:133-140): solves the session's graph and replies with the start / end
    time, the duration in seconds and `status`.  Upstream calls IIF's `solveTree!`; here the graph is a `solver.FactorGraph` and
    the solve is the caller harness `solver.solveGraph` on the device product path (`backend`: default `solver.DeviceBackend`)."""
    import datetime
    from . import solver
    if backend is None:
        import sys
        backend = solver.DeviceBackend(sys.modules[__package__])
    resp = {"startTime": datetime.datetime.now().isoformat()}
    solver.solveGraph(fg, backend, gibbs_iters=gibbs_iters, seed=seed)
    end = datetime.datetime.now()
    resp["endTime"] = end.isoformat()
    resp["durationSec"] = (end - datetime.datetime.fromisoformat(resp["startTime"])).total_seconds()
    resp["status"] = "OK"
    return resp
