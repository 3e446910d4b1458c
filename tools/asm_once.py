"""Dev tool for ncu: one warm-up and one profiled assembly of a tet cube or an icosphere.
Usage: python tools/asm_once.py tet 121 | tria 9"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M  # noqa: E402

kind = sys.argv[1]
size = int(sys.argv[2])
msh = M.cube_tets(size) if kind == "tet" else M.icosphere(size)
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
for _ in range(2):
    dm.drop_cache()
    a, b = _lib.assemble(ctx, dm, 3 if kind == "tet" else 0, False)
    del a, b
ctx.sync()
