"""Dev tool for the ncu launch list: two bench steps (cold assembly + eigs k=50 of the level-9 icosphere, mesh
resident in HBM); tools/launch_summary.py --after-last strip_rows_kernel<false cuts the list at the second step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lapy_b200 import _lib, mesh as M

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
mesh = M.icosphere(level)
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, mesh.v, mesh.t)
buf = np.zeros((mesh.v.shape[0], 50))
for i in range(2):
    dm.drop_cache()
    a, b = _lib.assemble(ctx, dm, _lib.FEM_TRIA, False)
    ev, evec, info = _lib.eigs(ctx, a, b, 50, -0.01, out_evecs=buf)
    print("step", i, info, flush=True)
