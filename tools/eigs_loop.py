"""Dev tool: wall clock and solver info of consecutive assemble + eigs(k=50) steps on one mesh."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lapy_b200 import _lib, mesh as M

what = sys.argv[1] if len(sys.argv) > 1 else "ico9"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mesh = M.icosphere(int(what[3:])) if what.startswith("ico") else M.cube_tets(int(what[4:]))
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, mesh.v, mesh.t)
kind = _lib.FEM_TETRA if mesh.t.shape[1] == 4 else _lib.FEM_TRIA
buf = np.zeros((mesh.v.shape[0], 50))
for i in range(reps):
    t0 = time.perf_counter()
    dm.drop_cache()
    a, b = _lib.assemble(ctx, dm, kind, False)
    ctx.sync()
    t1 = time.perf_counter()
    ev, evec, info = _lib.eigs(ctx, a, b, 50, -0.01, out_evecs=buf)
    t2 = time.perf_counter()
    print(f"step {i}: assemble {1e3*(t1-t0):.1f} ms, eigs {1e3*(t2-t1):.1f} ms, info {info}", flush=True)
