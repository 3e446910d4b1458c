"""Oracle: COO -> canonical CSC with duplicate summation (TEST INFRASTRUCTURE).

The reference builds every matrix with ``scipy.sparse.csc_matrix((data,(i,j)))``
(lapy/solver.py:178, :187, :193, :293, :301, :307, :370, :376, :501, :526, :532).  SciPy
(``_coo.py`` tocsc -> ``coo_tocsr`` on the transposed triplets, then ``sum_duplicates``:
``csr_sort_indices`` + ``csr_sum_duplicates``, ``_compressed.py``) does:

1. stable counting sort of the triplets by column  -> per column, triplets in COO input order;
2. ``std::sort`` of each column's (row, value) pairs by row (libstdc++ introsort: identical to
   a stable insertion sort for <= 16 pairs, deterministic but not stable above that);
3. left-to-right summation of equal rows, starting from the first addend itself;
4. int32 index arrays when max(nnz_coo, n) < 2**31; shape inferred = max index + 1; explicit
   zeros are kept.

:func:`coo_to_csc_sequential` restates 1-4 with a *stable* sort in step 2, i.e. every
(row, col) sum runs over its addends in COO input order.  This is the summation order of the
CUDA assembly (DESIGN.md "assembly"), so kernel output must equal this function bit for bit,
and it equals SciPy bit for bit whenever a column holds <= 16 triplets or all permutations
of the addends round identically (always true for <= 2 addends).
"""

from __future__ import annotations

import numpy as np


def coo_to_csc_sequential(dat, i, j, n=None):
    """Return (indptr int32 (n+1,), indices int32 (nnz,), data (nnz,)) - see module doc."""
    dat = np.asarray(dat)
    i = np.asarray(i).astype(np.int64).reshape(-1)
    j = np.asarray(j).astype(np.int64).reshape(-1)
    if n is None:
        n = int(max(i.max(), j.max())) + 1 if i.size else 0
    order = np.lexsort((i, j))  # stable: primary j, secondary i, ties in input order
    si, sj, sd = i[order], j[order], dat[order]
    if si.size == 0:
        return np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, dat.dtype)
    head = np.ones(si.size, bool)
    head[1:] = (si[1:] != si[:-1]) | (sj[1:] != sj[:-1])
    seg = np.cumsum(head) - 1
    starts = np.flatnonzero(head)
    rank = np.arange(si.size) - starts[seg]
    out = sd[starts].copy()  # first addend itself (keeps -0.0)
    for r in range(1, int(rank.max()) + 1):
        m = rank == r
        out[seg[m]] = out[seg[m]] + sd[m]
    indices = si[starts].astype(np.int32)
    counts = np.bincount(sj[starts], minlength=n)
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(counts)
    return indptr, indices, out
