"""Independent NumPy / pure-Python mirror of oracle/cb2_oracle.c (TEST INFRASTRUCTURE ONLY).

Second restatement of the same algorithms, written separately so that a slip in the C
file shows up as a disagreement.  Pure-Python loops: small cases only.
"""
import math

import numpy as np

TWO_PI = 2.0 * math.pi
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(ctr, key):
    c0, c1, c2, c3 = [int(x) for x in ctr]
    k0, k1 = [int(x) for x in key]
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def _u(x0, x1, ubits):
    if ubits == 24:
        return ((x0 >> 9) + 0.5) / 8388608.0
    return ((((x0 << 32) | x1) >> 11) + 0.5) / 9007199254740992.0


def uniform(seed, sample, draw, stream, tag, ubits=53):
    o = philox4x32((sample, draw, stream, tag), (seed & _MASK, (seed >> 32) & _MASK))
    return _u(o[0], o[1], ubits)


def normal(seed, sample, i, stream, tag, ubits=53):
    o = philox4x32((sample, i >> 1, stream, tag), (seed & _MASK, (seed >> 32) & _MASK))
    u1, u2 = _u(o[0], o[1], ubits), _u(o[2], o[3], ubits)
    r = math.sqrt(-2.0 * math.log(u1))
    return r * (math.sin(TWO_PI * u2) if i & 1 else math.cos(TWO_PI * u2))


def wrap_pi(x):
    return x - TWO_PI * np.round(x / TWO_PI)


def pair_sum(p, q, bw):
    d2 = ((p[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    return np.exp(-bw * d2).sum()


def mmd(a, b, bw):
    N, M = len(a), len(b)
    return pair_sum(a, a, bw) / N**2 - 2 * pair_sum(a, b, bw) / (N * M) + pair_sum(b, b, bw) / M**2


def se_matrix(dim, c, backward=False):
    c = np.asarray(c, float)
    if dim == 2:
        R = np.array([[math.cos(c[2]), -math.sin(c[2])], [math.sin(c[2]), math.cos(c[2])]])
    else:
        from scipy.linalg import expm
        w = c[3:]
        R = expm(np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]))
    t = c[:dim].copy()
    if backward:
        R = R.T
        t = -R @ t
    return R, t


def kde_eval(mu, sigma, q, w=None):
    N = len(mu)
    w = np.full(N, 1.0 / N) if w is None else np.asarray(w, float) / np.sum(w)
    z = (q[:, None, :] - mu[None, :, :]) / sigma
    k = np.exp(-0.5 * (z**2).sum(-1)) / np.prod(np.sqrt(TWO_PI) * sigma)
    return k @ w


def loo_nll_1d(x, h):
    n = len(x)
    d = x[:, None] - x[None, :]
    K = np.exp(-0.5 * (d / h) ** 2) / (math.sqrt(TWO_PI) * h)
    np.fill_diagonal(K, 0)
    p = K.sum(1) / (n - 1)
    return -np.mean(np.log(np.maximum(p, 1e-300)))


def golden(f, ax, bx, cx, tol=1e-2):
    Cc = (3 - math.sqrt(5)) / 2
    R = 1 - Cc
    x0, x3 = ax, cx
    if abs(cx - bx) > abs(bx - ax):
        x1, x2 = bx, bx + Cc * (cx - bx)
    else:
        x2, x1 = bx, bx - Cc * (bx - ax)
    f1, f2 = f(x1), f(x2)
    while abs(x3 - x0) > tol * (abs(x1) + abs(x2)):
        if f2 < f1:
            x0, x1, x2 = x1, x2, R * x2 + Cc * x3
            f1, f2 = f2, f(x2)
        else:
            x3, x2, x1 = x2, x1, R * x1 + Cc * x0
            f2, f1 = f1, f(x1)
    return x1 if f1 < f2 else x2


def bw_loo_1d(x):
    xs = np.sort(x)
    mn = max(np.diff(xs).min(), 1e-6)
    return golden(lambda h: loo_nll_1d(x, h), 2 * mn / (1 + math.sqrt(5)), mn, 2 * (xs[-1] - xs[0]))


# ---------------------------------------------------------------- level hierarchy + Gibbs product
def build_tree(pts, bw, w=None):
    pts = np.asarray(pts, float)
    N, d = pts.shape
    w = np.full(N, 1.0 / N) if w is None else np.asarray(w, float) / np.sum(w)
    depth = 0
    while (1 << depth) < N:
        depth += 1
    perm = list(range(N))
    starts = [[0, N]]
    for l in range(depth):
        nxt = []
        for k in range(len(starts[l]) - 1):
            s, e = starts[l][k], starts[l][k + 1]
            n = e - s
            nxt.append(s)
            if n >= 2:
                seg = perm[s:e]
                rng = [pts[seg, i].max() - pts[seg, i].min() for i in range(d)]
                best = int(np.argmax(rng))  # first max
                order = sorted(range(n), key=lambda r: (pts[seg[r], best], r))
                perm[s:e] = [seg[r] for r in order]
                nxt.append(s + (n + 1) // 2)
        nxt.append(N)
        starts.append(nxt)
    mean = [None] * (depth + 1)
    var = [None] * (depth + 1)
    wt = [None] * (depth + 1)
    mean[depth] = pts[perm].copy()
    var[depth] = np.tile(np.asarray(bw, float) ** 2, (N, 1))
    wt[depth] = w[perm].copy()
    for l in range(depth - 1, -1, -1):
        cnt = len(starts[l]) - 1
        mean[l], var[l], wt[l] = np.zeros((cnt, d)), np.zeros((cnt, d)), np.zeros(cnt)
        for k in range(cnt):
            if l + 1 == depth:
                ch = list(range(starts[l][k], starts[l][k + 1]))
            else:
                ch = [2 * k, 2 * k + 1]
            ws = wt[l + 1][ch].sum()
            m = (wt[l + 1][ch, None] * mean[l + 1][ch]).sum(0) / ws
            v = (wt[l + 1][ch, None] * (var[l + 1][ch] + (mean[l + 1][ch] - m) ** 2)).sum(0) / ws
            mean[l][k], var[l][k], wt[l][k] = m, v, ws
    return dict(N=N, d=d, depth=depth, perm=np.array(perm), starts=starts, mean=mean, var=var, wt=wt)


def _combine(trees, lev, ind, skip, circ):
    d = trees[0]["d"]
    M, Cv = np.zeros(d), np.zeros(d)
    for i in range(d):
        lam = sx = cx = sm = 0.0
        for k, t in enumerate(trees):
            if k == skip:
                continue
            l = 1.0 / t["var"][lev[k]][ind[k], i]
            m = t["mean"][lev[k]][ind[k], i]
            lam += l
            if circ[i]:
                sx += l * math.sin(m)
                cx += l * math.cos(m)
            else:
                sm += l * m
        Cv[i] = 1.0 / lam
        M[i] = math.atan2(sx, cx) if circ[i] else sm / lam
    return M, Cv


def product(densities, Nout, seed, circular=None, niter=1, stream=0, ubits=53, level_down=1):
    trees = [build_tree(*dn) for dn in densities]
    D, d = len(trees), trees[0]["d"]
    circ = [0] * d if circular is None else list(circular)
    maxN = max(t["N"] for t in trees)
    L = int(math.floor(math.log2(maxN))) + 1
    sweeps = 1 + niter
    out = np.zeros((Nout, d))
    labels = np.zeros((Nout, D), dtype=np.int32)
    for s in range(Nout):
        ind, lev = [0] * D, [0] * D
        for step in range(1, L + 1):
            for j, t in enumerate(trees):
                if lev[j] < t["depth"]:  # levelDown: a child of the current node, drawn in proportion to the children's weights
                    lc = lev[j] + 1
                    if lc == t["depth"]:
                        c0 = t["starts"][lev[j]][ind[j]]
                        nch = t["starts"][lev[j]][ind[j] + 1] - c0
                    else:
                        c0, nch = 2 * ind[j], 2
                    pick = c0
                    if nch > 1 and level_down:
                        w0, w1 = t["wt"][lc][c0], t["wt"][lc][c0 + 1]
                        u = uniform(seed, s, (step - 1) * D + j, stream, 2, ubits)
                        if not (w0 >= u * (w0 + w1)):
                            pick = c0 + 1
                    ind[j] = pick
                    lev[j] = lc
            for it in range(sweeps):
                for j, t in enumerate(trees):
                    mu, va, wt = t["mean"][lev[j]], t["var"][lev[j]], t["wt"][lev[j]]
                    e = np.log(wt)
                    if D > 1:
                        M, Cv = _combine(trees, lev, ind, j, circ)
                        df = mu - M
                        for i in range(d):
                            if circ[i]:
                                df[:, i] = wrap_pi(df[:, i])
                        sv = va + Cv
                        e = e - 0.5 * (df**2 / sv + np.log(sv)).sum(1)
                    p = np.exp(e - e.max())
                    cs = np.cumsum(p)
                    u = uniform(seed, s, ((step - 1) * sweeps + it) * D + j, stream, 0, ubits)
                    z = int(np.searchsorted(cs, u * cs[-1], side="left"))
                    ind[j] = min(z, len(p) - 1)
        labels[s] = [t["perm"][ind[j]] for j, t in enumerate(trees)]
        M, Cv = _combine(trees, lev, ind, -1, circ)
        for i in range(d):
            v = M[i] + math.sqrt(Cv[i]) * normal(seed, s, i, stream, 1, ubits)
            out[s, i] = wrap_pi(v) if circ[i] else v
    return out, labels
