import os, sys, time
os.environ["LAPY_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 9
m = M.icosphere(lvl)
ctx = _lib.Context(0)
dm = _lib.DeviceMesh(ctx, m.v, m.t)
for it in range(4):
    print("--- iteration", it, file=sys.stderr)
    dm.drop_cache()
    t0 = time.perf_counter()
    a, b = _lib.assemble(ctx, dm, _lib.FEM_TRIA, False)
    print("total wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
