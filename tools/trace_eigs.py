"""Development aid: ShapeDNA k=50 on an icosphere / tet cube with the library's phase trace."""
import os, sys, time
os.environ.setdefault("LAPY_B200_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lapy_b200
from lapy_b200 import mesh as M

what = sys.argv[1] if len(sys.argv) > 1 else "ico7"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t0 = time.perf_counter()
mesh = M.icosphere(int(what[3:])) if what.startswith("ico") else M.cube_tets(int(what[4:]))
t1 = time.perf_counter()
fem = lapy_b200.Solver(mesh)
t2 = time.perf_counter()
ev, evec = fem.eigs(k=k)
t3 = time.perf_counter()
print(f"{what}: mesh {t1-t0:.2f}s  Solver {t2-t1:.3f}s  eigs {t3-t2:.3f}s  info {fem.last_info}")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "spectra.npz"))
key = f"{what}_k50"
if key in g and k == 50:
    ref = g[key]
    print("max rel err vs reference (1:)", np.max(np.abs(ev[1:] - ref[1:]) / ref[1:]), " lam0", ev[0], ref[0])
t4 = time.perf_counter()
ev2, _ = fem.eigs(k=k)
print(f"second eigs call {time.perf_counter()-t4:.3f}s  info {fem.last_info}")
