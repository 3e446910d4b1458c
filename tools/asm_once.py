"""Dev tool: a few cold assemblies of one mesh (for LAPY_B200_TRACE=1 / ncu).  Usage: asm_once.py [ico9|cube121] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M
what = sys.argv[1] if len(sys.argv) > 1 else "ico9"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
msh = M.icosphere(int(what[3:])) if what.startswith("ico") else M.cube_tets(int(what[4:]))
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
for _ in range(reps):
    dm.drop_cache()
    ctx.timer_start()
    a, b = _lib.assemble(ctx, dm, 3 if msh.t.shape[1] == 4 else 0, False)
    print("assemble ms", ctx.timer_stop(), flush=True)
    del a, b
