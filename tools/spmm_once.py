"""Dev tool: a few launches of the dominant SpMM (level-9 stiffness x 64-column block, solver numbering) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M
m = int(sys.argv[1]) if len(sys.argv) > 1 else 64
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # 2: the single-precision strip kernel
msh = M.icosphere(9)
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
a, b = _lib.assemble(ctx, dm, 0, False)
print("ms per launch", _lib.spmm_benchmark(ctx, a, m, 3, renumber=True, variant=variant))
