"""Dev tool: device time and roofline fraction of y = K x for every block width the solvers use.

    python tools/spmm_shapes.py [level] [cube_n]   -> gpurun_out/spmm_shapes.json

K = stiffness of the level-`level` icosphere / n^3 tet cube in the solver (Morton-cell) numbering;
x, y resident in HBM; CUDA events over 20 launches per width.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
cube_n = int(sys.argv[2]) if len(sys.argv) > 2 else 121
PEAK = 6550.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = _lib.default_context()
out = {}
for name, msh, kind in (("ico%d" % level, M.icosphere(level), 0), ("cube%d" % cube_n, M.cube_tets(cube_n), 3)):
    dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
    a, b = _lib.assemble(ctx, dm, kind, False)
    for m in (1, 2, 4, 8, 12, 16, 20, 24, 32, 40, 48, 56, 64, 96, 128):
        nbytes = 12.0 * a.nnz + 4.0 * (a.n + 1) + 16.0 * a.n * m
        for ren, label in ((1, "solver-order"), (0, "caller-order")):
            if ren == 0 and m not in (1, 16, 64):
                continue
            ms = _lib.spmm_benchmark(ctx, a, m, 20, renumber=ren)
            out[f"{name} m={m} {label}"] = {"ms": ms, "gb_s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK}
            print(name, m, label, f"{ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s  {nbytes / ms / 1e6 / PEAK:.3f}", flush=True)
    del a, b, dm
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/spmm_shapes.json", "w"), indent=1)
