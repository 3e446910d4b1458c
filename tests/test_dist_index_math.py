"""Index arithmetic of the two opt-in communication schemes of the row-partitioned mode, emulated in
NumPy for W ranks in one process (CPU only).  The formulas mirror csrc/eigs.cu one to one:

* ``dist_precond``: offsets / counts (so, sc, ro, rc) of the all-to-all that turns the row-partitioned
  residual block (n_loc, ma) into the column-partitioned (n, ma/W) one and back;
* ``build_halo`` + the halo SpMM: ghost lists grouped by owner, request exchange, the row block with
  columns renumbered to [own | ghosts].

``exchange`` has the semantics of ``dist_exchange`` (grouped ncclSend / ncclRecv).  The emulation checks
that matching send / receive counts agree on both sides, that the round trip reproduces a column-wise
linear map applied to the full block, and that the halo SpMM equals the global SpMM.
"""

import numpy as np
import scipy.sparse as sp


def exchange(W, sends, off_s, cnt_s, recvs, off_r, cnt_r):
    for r in range(W):
        for p in range(W):
            if p == r:
                continue
            if cnt_s[r][p] > 0:
                seg = sends[r][off_s[r][p] : off_s[r][p] + cnt_s[r][p]]
                assert cnt_r[p][r] == cnt_s[r][p], (r, p, cnt_r[p][r], cnt_s[r][p])
                recvs[p][off_r[p][r] : off_r[p][r] + cnt_r[p][r]] = seg


import pytest


@pytest.mark.parametrize("W,n,ma", [(2, 1001, 64), (3, 1000, 5), (4, 997, 1), (8, 4099, 37)])
def test_column_parallel_preconditioner_round_trip(W, n, ma):
    rng = np.random.default_rng(W * 1000 + ma)
    rpr = (n + W - 1) // W

    def row0(p):
        return min(n, p * rpr)

    def rows(p):
        return min(n, (p + 1) * rpr) - row0(p)

    mc = (ma + W - 1) // W

    def c0(p):
        return min(ma, p * mc)

    def cw(p):
        return min(ma, (p + 1) * mc) - c0(p)

    R = rng.standard_normal((n, ma))
    M = rng.standard_normal((n, n)) / n  # stands for the (linear, column-wise) AMG cycle
    tpack = [np.zeros(max(1, rows(r) * ma)) for r in range(W)]
    tfr = [np.zeros(n * mc) for _ in range(W)]
    tfz = [np.zeros(n * mc) for _ in range(W)]
    so, sc, ro, rc = ([[0] * W for _ in range(W)] for _ in range(4))
    for me in range(W):
        nl, mine, r = rows(me), cw(me), R[row0(me) : row0(me) + rows(me)]
        for p in range(W):
            so[me][p], sc[me][p] = nl * c0(p), nl * cw(p)
            ro[me][p], rc[me][p] = row0(p) * mine, rows(p) * mine
            if cw(p) == 0:
                continue
            blk = r[:, c0(p) : c0(p) + cw(p)].reshape(-1)
            if p == me:
                tfr[me][ro[me][p] : ro[me][p] + nl * mine] = blk
            else:
                tpack[me][so[me][p] : so[me][p] + nl * cw(p)] = blk
    exchange(W, tpack, so, sc, tfr, ro, rc)
    for me in range(W):
        mine = cw(me)
        if mine:
            X = tfr[me][: n * mine].reshape(n, mine)
            np.testing.assert_array_equal(X, R[:, c0(me) : c0(me) + mine])
            tfz[me][: n * mine] = (M @ X).reshape(-1)
    back = [np.zeros(max(1, rows(r) * ma)) for r in range(W)]
    exchange(W, tfz, ro, rc, back, so, sc)
    Z = np.zeros((n, ma))
    for me in range(W):
        nl, mine = rows(me), cw(me)
        for p in range(W):
            if cw(p) == 0:
                continue
            if p == me:
                blk = tfz[me][ro[me][p] : ro[me][p] + nl * mine].reshape(nl, mine)
            else:
                blk = back[me][so[me][p] : so[me][p] + nl * cw(p)].reshape(nl, cw(p))
            Z[row0(me) : row0(me) + nl, c0(p) : c0(p) + cw(p)] = blk
    np.testing.assert_allclose(Z, M @ R, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("W,n", [(2, 500), (3, 777), (4, 1000)])
def test_halo_plan_spmm_equals_global_spmm(W, n):
    rng = np.random.default_rng(n)
    A = sp.random(n, n, density=0.01, random_state=1, format="csr")
    A = (A + A.T + sp.eye(n)).tocsr()
    A.sort_indices()
    rpr, w = (n + W - 1) // W, 3
    X = rng.standard_normal((n, w))
    Yref = A @ X
    plans = []
    for me in range(W):
        r0 = min(n, me * rpr)
        r1 = min(n, r0 + rpr)
        blk = A[r0:r1]
        idx = blk.indices
        ghosts = np.unique(idx[(idx < r0) | (idx >= r1)])
        recv_cnt = [int(((ghosts // rpr) == p).sum()) for p in range(W)]
        recv_off = np.concatenate([[0], np.cumsum(recv_cnt)[:-1]]).astype(int).tolist()
        plans.append(dict(r0=r0, r1=r1, blk=blk, ghosts=ghosts, recv_cnt=recv_cnt, recv_off=recv_off))
    for me, P in enumerate(plans):
        P["send_cnt"] = [0 if p == me else plans[p]["recv_cnt"][me] for p in range(W)]  # the all-gathered counts
        P["send_off"] = np.concatenate([[0], np.cumsum(P["send_cnt"])[:-1]]).astype(int).tolist()
        P["n_send"] = sum(P["send_cnt"])
    reqs = [np.zeros(max(1, P["n_send"]), np.int64) for P in plans]
    exchange(W, [P["ghosts"].astype(np.int64) for P in plans], [P["recv_off"] for P in plans],
             [P["recv_cnt"] for P in plans], reqs, [P["send_off"] for P in plans], [P["send_cnt"] for P in plans])  # fmt: skip
    for me, P in enumerate(plans):
        P["send_rows"] = reqs[me][: P["n_send"]] - P["r0"]
        assert ((P["send_rows"] >= 0) & (P["send_rows"] < P["r1"] - P["r0"])).all()
    xg, sb = [], []
    for P in plans:
        nl, ng = P["r1"] - P["r0"], len(P["ghosts"])
        g = np.zeros((nl + ng) * w)
        g[: nl * w] = X[P["r0"] : P["r1"]].reshape(-1)
        xg.append(g)
        sb.append(X[P["r0"] : P["r1"]][P["send_rows"]].reshape(-1) if P["n_send"] else np.zeros(1))
    exchange(W, sb, [[o * w for o in P["send_off"]] for P in plans], [[c * w for c in P["send_cnt"]] for P in plans], xg,
             [[(P["r1"] - P["r0"] + o) * w for o in P["recv_off"]] for P in plans],
             [[c * w for c in P["recv_cnt"]] for P in plans])  # fmt: skip
    for me, P in enumerate(plans):
        nl, idx = P["r1"] - P["r0"], P["blk"].indices
        loc = np.where((idx >= P["r0"]) & (idx < P["r1"]), idx - P["r0"], nl + np.searchsorted(P["ghosts"], idx))
        L = sp.csr_matrix((P["blk"].data, loc, P["blk"].indptr), shape=(nl, nl + len(P["ghosts"])))
        np.testing.assert_allclose(L @ xg[me].reshape(-1, w), Yref[P["r0"] : P["r1"]], rtol=1e-12, atol=1e-12)
