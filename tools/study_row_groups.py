"""CPU study for the row-grouped SpMM idea: how many distinct X rows does a group of
R consecutive matrix rows touch, under (a) the current solver numbering (Morton order of a 128^3 cell
grid, original index inside a cell) and (b) a vertex-granular Morton order (2^20 grid)?
Loads per row = union size / R; the CSR kernel does nnz/row (7 on a closed triangle mesh).

    python tools/study_row_groups.py [icosphere level | cubeN]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import mesh as M  # noqa: E402


def spread3(x):
    x = x.astype(np.uint64)
    out = np.zeros_like(x)
    for b in range(21):
        out |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out


def morton(v, bits):
    lo, hi = v.min(0), v.max(0)
    g = (1 << bits) - 1
    q = np.clip(((v - lo) / np.maximum(hi - lo, 1e-300) * (g + 1)).astype(np.int64), 0, g)
    return spread3(q[:, 0]) | (spread3(q[:, 1]) << np.uint64(1)) | (spread3(q[:, 2]) << np.uint64(2))


arg = sys.argv[1] if len(sys.argv) > 1 else "7"
msh = M.cube_tets(int(arg[4:])) if arg.startswith("cube") else M.icosphere(int(arg))
n, k = len(msh.v), msh.t.shape[1]
i = np.repeat(msh.t, k, axis=1).reshape(-1)
j = np.tile(msh.t, (1, k)).reshape(-1)
A = sp.csr_matrix((np.ones(len(i)), (i, j)), shape=(n, n))
A.sum_duplicates()
print(f"{arg}: n={n}, nnz/row={A.nnz / n:.2f}")
orders = {
    "caller": np.arange(n),
    "cells128 (current)": np.lexsort((np.arange(n), morton(msh.v, 7))),
    "morton 2^20 (vertex granular)": np.argsort(morton(msh.v, 20), kind="stable"),
}
for name, order in orders.items():
    inv = np.empty(n, np.int64)
    inv[order] = np.arange(n)
    P = A[order][:, order].tocsr()
    P.sort_indices()
    line = [f"{name:32s}"]
    for R in (1, 2, 4, 8):
        ng = (n + R - 1) // R
        grp = np.repeat(np.arange(n) // R, np.diff(P.indptr))
        key = grp.astype(np.int64) * n + P.indices
        union = np.unique(key).size
        line.append(f"R={R}: {union / n:5.2f} loads/row")
    # strip locality: fraction of nonzeros whose column lies in the row's own 128-row strip
    rows = np.repeat(np.arange(n), np.diff(P.indptr))
    line.append(f"in-strip(128) {np.mean(rows // 128 == P.indices // 128):.2f}")
    print("  ".join(line))
