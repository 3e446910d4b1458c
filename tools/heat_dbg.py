import sys, time, os
os.environ["LAPY_B200_TRACE"]="1"
sys.path.insert(0, ".")
import numpy as np
from lapy_b200 import mesh as M, heat
for lvl in (6, 8, 9):
    m = M.icosphere(lvl)
    t0=time.perf_counter(); u = heat.diffusion(m, [0]); print("L",lvl,"heat wall",time.perf_counter()-t0, heat.diffusion.last_info, "min",u.min(),"max",u.max(), file=sys.stderr)
