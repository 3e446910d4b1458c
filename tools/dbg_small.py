import os, sys, time
os.environ["LAPY_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lapy_b200
from lapy_b200 import mesh as M
for name, mk, k in (("ico3", lambda: M.icosphere(3), 20), ("ico5", lambda: M.icosphere(5), 50), ("cube9", lambda: M.cube_tets(9), 12), ("ico6", lambda: M.icosphere(6), 50)):
    m = mk(); fem = lapy_b200.Solver(m)
    t0 = time.perf_counter(); ev, _ = fem.eigs(k=k); t1 = time.perf_counter()
    print(f"== {name} k={k}: {t1-t0:.3f}s {fem.last_info}", file=sys.stderr)
