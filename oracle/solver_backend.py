"""CPU-oracle backend for caesar.jl_b200/solver.py (TEST INFRASTRUCTURE: the CPU arm of the hex-solve comparison)."""
import numpy as np

import oracle

TOPO = {"Pose2": [0, 0, 1], "Point2": [0, 0], "ContinuousScalar": [0]}


class _Belief:
    def __init__(self, vartype, pts, bw):
        self.manifold, self.pts, self.bw = vartype, np.ascontiguousarray(pts, dtype=np.float64), np.asarray(bw, dtype=np.float64)
        self.N, self.d = self.pts.shape
        self.circular = tuple(TOPO[vartype])


class OracleBackend:
    def __init__(self, ubits=53):
        self.ubits = ubits

    def manikde_batch(self, vartype, pts_list):
        return [_Belief(vartype, p, oracle.kde_bw_loo(p, circular=TOPO[vartype])) for p in pts_list]

    def products(self, vartype, groups, N, seed, streams):
        out = []
        for g, st in zip(groups, streams):
            if len(g) == 1:
                out.append(g[0])
                continue
            pts, _ = oracle.product([(m.pts, m.bw) for m in g], N, seed, circular=TOPO[vartype], stream=st, ubits=self.ubits)
            out.append(_Belief(vartype, pts, oracle.kde_bw_loo(pts, circular=TOPO[vartype])))
        return out

    def approx_conv_batch(self, convs, seed):
        return [oracle.approx_conv(c["kind"], c["reverse"], c["N"], c["mean"], c["chol"], seed, stream=c["stream"], src=c["src"], aux=c["aux"])
                for c in convs]
