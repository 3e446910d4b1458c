import os, sys, time
os.environ["LAPY_B200_TRACE"] = "1"; os.environ["LAPY_B200_FORCE_DIST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lapy_b200
from lapy_b200 import _lib, mesh as M
ctx = _lib.Context(0); ctx.init_row_partition_single()
mesh = M.icosphere(6)
fem = lapy_b200.Solver(mesh, ctx=ctx)
try:
    ev, evec = fem.eigs(k=50, maxit=30)
    print("OK", fem.last_info, ev[:4])
except Exception as e:
    print("FAIL", e)
