"""Dev tool for compute-sanitizer: small instances of the hot path (assembly tri / tet, SpMM self-test, eigs, heat)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lapy_b200
from lapy_b200 import _lib, mesh as M, heat

ctx = _lib.default_context()
tri = M.icosphere(4)
dm = _lib.DeviceMesh(ctx, tri.v, tri.t)
a, b = _lib.assemble(ctx, dm, _lib.FEM_TRIA, False)
a2, b2 = _lib.assemble(ctx, dm, _lib.FEM_TRIA, True)
print("tri nnz", a.nnz, b2.nnz)
print("selftest", _lib.spmm_selftest(ctx, a, 24))
tet = M.cube_tets(7)
dmt = _lib.DeviceMesh(ctx, tet.v, tet.t)
at, bt = _lib.assemble(ctx, dmt, _lib.FEM_TETRA, False)
print("selftest tet", _lib.spmm_selftest(ctx, at, 8))
ev, evec = lapy_b200.Solver(M.icosphere(5)).eigs(k=12)
print("eigs", ev[:4])
ev2, _ = lapy_b200.Solver(M.icosphere(5)).eigs(k=12, vectors=False)
assert np.allclose(ev, ev2)
u = heat.diffusion(tri, [0])
print("heat", float(u[0]))
