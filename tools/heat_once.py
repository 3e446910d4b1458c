"""Dev tool: config 4 (heat diffusion + heat-method geodesics) on an icosphere, step by step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lapy_b200 import heat, diffgeo, mesh as M
level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
mesh = M.icosphere(level)
for m in (1.0, 16.0):
    t0 = time.perf_counter()
    u = heat.diffusion(mesh, [0], m=m)
    print("heat m=%g" % m, time.perf_counter() - t0, dict(heat.diffusion.last_info), float(u[0]), flush=True)
t0 = time.perf_counter()
g = diffgeo.compute_geodesic_f(mesh, u)
print("geodesic", time.perf_counter() - t0, float(g.max()), flush=True)
