"""Dev tool: lumped vs full assembly of one icosphere - A must be identical, the lumped mass = row sums of the full mass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lapy_b200
from lapy_b200 import mesh as M
level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
mesh = M.icosphere(level)
for rep in range(3):
    f = lapy_b200.Solver(mesh, lump=False)
    a0, b0 = f.stiffness, f.mass
    g = lapy_b200.Solver(mesh, lump=True)
    a1, b1 = g.stiffness, g.mass
    same = np.array_equal(a0.indptr, a1.indptr) and np.array_equal(a0.indices, a1.indices) and np.array_equal(a0.data, a1.data)
    print("rep", rep, "A identical", same, "indices max", a1.indices.max(), "n", a1.shape[0], "nnz", a0.nnz, a1.nnz,
          "indptr ok", bool(np.all(np.diff(a1.indptr) >= 0)), a1.indptr[-1],
          "lump vs rowsum", float(np.abs(b1.diagonal() - np.asarray(b0.sum(axis=1)).ravel()).max()), flush=True)
    if not same:
        bad = np.nonzero(a0.indptr != a1.indptr)[0]
        print("  first differing indptr positions", bad[:10], a0.indptr[bad[:5]], a1.indptr[bad[:5]])
