"""CPU prototype (NumPy/SciPy) of the device solver algorithms - development aid only.

Mirrors, step by step, what lapy_b200/csrc/{amg,pcg,lobpcg}.cu do on the GPU so that algorithmic
parameters (aggregation, smoother, block size, locking) can be tuned without GPU time:
  * MIS-2 aggregation with deterministic hash priorities, smoothed-aggregation prolongator,
    Galerkin coarse operators, Chebyshev / Jacobi smoother, dense coarse solve;
  * block LOBPCG on (A, M) with explicit M-orthonormal basis [X, W, P] and soft locking;
  * block PCG with constant-null-space projection.
Not imported by the product, tests or bench.
"""

from __future__ import annotations

import sys
import time

import numpy as np
from scipy import sparse


def hash32(i):
    x = (i.astype(np.uint64) + 1) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(32)
    return (x & np.uint64(0x7FFFFFFF)).astype(np.int64)


def mis2_aggregate(S):
    """S: symmetric strength graph (csr, no diagonal needed). Returns agg ids (n,), n_agg."""
    n = S.shape[0]
    S = S.tocsr()
    rows = np.repeat(np.arange(n), np.diff(S.indptr))
    cols = S.indices
    prio = hash32(np.arange(n)) * n + np.arange(n)  # unique
    state = np.ones(n, np.int64)  # 1 undecided, 2 in, 0 out
    big = np.int64(n) * (1 << 31)
    rounds = 0
    while (state == 1).any():
        rounds += 1
        t = state * big * 4 + prio  # lexicographic (state, prio)
        t0 = t.copy()
        for _ in range(2):
            tn = t.copy()
            np.maximum.at(tn, rows, t[cols])
            t = tn
        und = state == 1
        win = und & (t == t0)
        lose = und & (t // (big * 4) == 2)
        state[win] = 2
        state[lose & ~win] = 0
    roots = np.flatnonzero(state == 2)
    agg = -np.ones(n, np.int64)
    agg[roots] = np.arange(len(roots))
    # pass 1: neighbours of roots (pick root with max priority)
    best = -np.ones(n, np.int64)
    isroot = state[cols] == 2
    np.maximum.at(best, rows[isroot], prio[cols[isroot]])
    m1 = (agg < 0) & (best >= 0)
    agg[m1] = agg[(best[m1] % n)]
    # pass 2: distance-2 nodes join the aggregate of the neighbour with max priority among assigned-in-pass-1/roots
    assigned = agg >= 0
    best2 = -np.ones(n, np.int64)
    ok = assigned[cols]
    np.maximum.at(best2, rows[ok], prio[cols[ok]])
    m2 = (agg < 0) & (best2 >= 0)
    agg[m2] = agg[(best2[m2] % n)]
    assert (agg >= 0).all(), "isolated nodes?"
    return agg, len(roots), rounds


def strength(K, theta):
    K = K.tocsr()
    d = K.diagonal()
    n = K.shape[0]
    rows = np.repeat(np.arange(n), np.diff(K.indptr))
    keep = (np.abs(K.data) >= theta * np.sqrt(np.abs(d[rows] * d[K.indices]))) & (rows != K.indices)
    return sparse.csr_matrix((np.ones(keep.sum()), (rows[keep], K.indices[keep])), shape=K.shape)


class AMG:
    def __init__(self, K, theta=0.0, max_coarse=1500, max_levels=12, cheb_deg=2, omega_scale=4.0 / 3.0, verbose=True):
        self.levels = []
        self.cheb_deg = cheb_deg
        K = K.tocsr()
        while True:
            n = K.shape[0]
            d = K.diagonal()
            dinv = 1.0 / d
            # spectral radius of D^-1 K by a few power iterations (deterministic start)
            x = np.cos(np.arange(n) * 0.7) + 1.5
            for _ in range(12):
                y = dinv * (K @ x)
                rho = np.linalg.norm(y) / np.linalg.norm(x)
                x = y / np.linalg.norm(y)
            rho *= 1.05
            lvl = {"K": K, "dinv": dinv, "rho": rho}
            self.levels.append(lvl)
            if n <= max_coarse or len(self.levels) >= max_levels:
                lvl["dense"] = np.linalg.inv(K.toarray())
                break
            S = strength(K, theta)
            agg, na, rounds = mis2_aggregate(S)
            cnt = np.bincount(agg, minlength=na)
            T = sparse.csr_matrix((1.0 / np.sqrt(cnt[agg]), (np.arange(n), agg)), shape=(n, na))
            omega = omega_scale / rho
            P = T - omega * sparse.diags(dinv) @ (K @ T)
            P = P.tocsr()
            lvl["P"] = P
            lvl["R"] = P.T.tocsr()
            K = (lvl["R"] @ K @ P).tocsr()
            K.sort_indices()
            if verbose:
                print(f"  level {len(self.levels)-1}: n={n} nnz={lvl['K'].nnz} -> n_c={na} (mis rounds {rounds}), P nnz/row {P.nnz/n:.2f}")
        if verbose:
            tot = sum(l["K"].nnz for l in self.levels)
            print(f"  levels={len(self.levels)} operator complexity={tot/self.levels[0]['K'].nnz:.3f} coarsest n={self.levels[-1]['K'].shape[0]}")

    def smooth(self, lvl, x, b, zero_guess):
        K, dinv, rho = lvl["K"], lvl["dinv"], lvl["rho"]
        deg = self.cheb_deg
        if deg == 0:  # weighted Jacobi
            w = 4.0 / (3.0 * rho)
            r = b if zero_guess else b - K @ x
            return (0 if zero_guess else x) + w * dinv[:, None] * r
        # Chebyshev on D^-1 K over [rho/alpha, rho]
        lo, hi = rho / 8.0, rho
        theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
        r = b.copy() if zero_guess else b - K @ x
        x = np.zeros_like(b) if zero_guess else x.copy()
        sigma = theta / delta
        rho_k = 1.0 / sigma
        dvec = dinv[:, None] * r / theta
        for k in range(deg):
            x = x + dvec
            if k == deg - 1:
                break
            r = r - K @ dvec
            rho_n = 1.0 / (2 * sigma - rho_k)
            dvec = rho_n * rho_k * dvec + 2 * rho_n / delta * dinv[:, None] * r
            rho_k = rho_n
        return x

    def vcycle(self, b, l=0):
        lvl = self.levels[l]
        if "dense" in lvl:
            return lvl["dense"] @ b
        x = self.smooth(lvl, None, b, True)
        r = b - lvl["K"] @ x
        xc = self.vcycle(lvl["R"] @ r, l + 1)
        x = x + lvl["P"] @ xc
        x = self.smooth(lvl, x, b, False)
        return x

    def __call__(self, b):
        return self.vcycle(b)


def m_orthonormalize(W, MW_fn):
    """Cholesky-QR in the M inner product, twice. Returns W, MW."""
    for _ in range(2):
        MW = MW_fn(W)
        G = W.T @ MW
        G = 0.5 * (G + G.T)
        try:
            L = np.linalg.cholesky(G)
            W = np.linalg.solve(L, W.T).T
        except np.linalg.LinAlgError:
            # SVQB fallback
            d = np.sqrt(np.diag(G))
            Gs = G / np.outer(d, d)
            ew, ev = np.linalg.eigh(Gs)
            ew = np.maximum(ew, 1e-14 * ew.max())
            W = (W / d) @ (ev / np.sqrt(ew))
    return W, MW_fn(W)


def lobpcg(A, M, prec, k, m=None, tol=1e-9, maxit=100, verbose=True, seed=0):
    n = A.shape[0]
    m = m or k + max(8, k // 4)
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, m))
    X[:, 0] = 1.0
    Mf = lambda V: M @ V  # noqa: E731
    X, MX = m_orthonormalize(X, Mf)
    AX = A @ X
    G = X.T @ AX
    lam, C = np.linalg.eigh(0.5 * (G + G.T))
    X, AX, MX = X @ C, AX @ C, MX @ C
    P = None
    hist = []
    for it in range(maxit):
        R = AX - MX * lam
        scale = np.maximum(np.abs(lam), np.abs(lam[:k]).mean())
        rn = np.linalg.norm(R, axis=0) / (scale * np.linalg.norm(MX, axis=0))
        conv = rn < tol
        nconv_k = int(conv[:k].sum())
        hist.append(rn[:k].max())
        if verbose:
            print(f"  it {it:3d}: max res (first k) {rn[:k].max():.3e}  converged {nconv_k}/{k}  active {int((~conv).sum())}")
        if conv[:k].all():
            break
        act = ~conv
        act[k:] = act[k:] & (np.arange(k, m) < k + (m - k))  # guards always active unless converged
        W = prec(R[:, act])
        # orthogonalise W against X (and P) in the M inner product, twice
        for _ in range(2):
            W = W - X @ (MX.T @ W)
            if P is not None:
                W = W - P @ (MP.T @ W)
        W, MW = m_orthonormalize(W, Mf)
        AW = A @ W
        if P is not None:
            S, AS = np.hstack([X, W, P]), np.hstack([AX, AW, AP])
        else:
            S, AS = np.hstack([X, W]), np.hstack([AX, AW])
        G = S.T @ AS
        G = 0.5 * (G + G.T)
        ew, C = np.linalg.eigh(G)
        Cx = C[:, :m]
        lam = ew[:m]
        # new search directions in coefficient space: drop the X part, orthonormalise against Cx
        Cp = Cx.copy()
        Cp[:m, :] = 0.0
        Cp = Cp[:, act] if True else Cp
        Cp = Cp - Cx @ (Cx.T @ Cp)
        Cp = Cp - Cx @ (Cx.T @ Cp)
        q, r = np.linalg.qr(Cp)
        Cp = q
        MS = np.hstack([MX, MW, MP]) if P is not None else np.hstack([MX, MW])
        X, AX, MX = S @ Cx, AS @ Cx, MS @ Cx
        P, AP, MP = S @ Cp, AS @ Cp, MS @ Cp
    return lam[:k], X[:, :k], it + 1, hist


def pcg(Kop, b, prec, tol=1e-10, maxit=500, project=False, verbose=False):
    x = np.zeros_like(b)
    if project:
        b = b - b.mean(0)
    r = b.copy()
    z = prec(r)
    if project:
        z = z - z.mean(0)
    p = z.copy()
    rz = (r * z).sum(0)
    b0 = np.linalg.norm(b, axis=0)
    for it in range(maxit):
        Kp = Kop(p)
        alpha = rz / (p * Kp).sum(0)
        x += alpha * p
        r -= alpha * Kp
        rn = np.linalg.norm(r, axis=0) / b0
        if verbose:
            print(f"   pcg {it}: {rn.max():.3e}")
        if rn.max() < tol:
            break
        z = prec(r)
        if project:
            z = z - z.mean(0)
        rz_new = (r * z).sum(0)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it + 1


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from lapy_b200 import mesh as Mh
    from oracle import fem as ofem

    what = sys.argv[1] if len(sys.argv) > 1 else "ico6"
    if what.startswith("ico"):
        mesh = Mh.icosphere(int(what[3:]))
    else:
        mesh = Mh.cube_tets(int(what[4:]))
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    A, B = ofem.fem(mesh)
    A, B = A.tocsr(), B.tocsr()
    K = (A + 0.01 * B).tocsr()
    t0 = time.time()
    amg = AMG(K, cheb_deg=int(sys.argv[3]) if len(sys.argv) > 3 else 2)
    print("amg setup", time.time() - t0)
    t0 = time.time()
    lam, X, its, hist = lobpcg(A, B, amg, k, m=int(sys.argv[4]) if len(sys.argv) > 4 else None)
    print("lobpcg its", its, "time", time.time() - t0)
    g = np.load("tests/golden/spectra.npz")
    key = f"{what}_k50"
    if key in g and k == 50:
        ref = g[key]
        print("max rel err vs golden (1:)", np.max(np.abs(lam[1:] - ref[1:]) / ref[1:]), "lam0", lam[0], ref[0])
