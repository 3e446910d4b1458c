"""Dev tool: device time of lb_fem_assemble (cold incidence cache) on the level-L icosphere and the
n^3 tet cube.  Usage (GPU box): python tools/time_assembly.py [level] [cube_n]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
cube_n = int(sys.argv[2]) if len(sys.argv) > 2 else 121
ctx = _lib.default_context()
out = {}
for name, msh, kind in (("tria", M.icosphere(level), 0), ("tet", M.cube_tets(cube_n), 3)):
    dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
    for lump in (False, True):
        ms = []
        for _ in range(8):
            dm.drop_cache()
            ctx.timer_start()
            a, b = _lib.assemble(ctx, dm, kind, lump)
            ms.append(ctx.timer_stop())
        nnz = a.nnz
        del a, b
        nt, nv, k = len(msh.t), len(msh.v), msh.t.shape[1]
        nbytes = 4 * k * nt + 24 * nv + (12 * nnz + 4 * (nv + 1)) * (1 if lump else 2) + (12 * nv + 4 * (nv + 1) if lump else 0)
        best = min(ms[2:])
        out[f"{name}_lump{int(lump)}"] = {"nt": nt, "nv": nv, "nnz": nnz, "ms_best": best, "ms_median": float(np.median(ms[2:])),
                                          "gelem_s": nt / best / 1e6, "roofline_frac": nbytes / best / 1e6 / 6550.1}
        print(name, lump, out[f"{name}_lump{int(lump)}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_assembly.json", "w"), indent=1)
