"""Dev tool: SpMV / SpMM-64 kernel time and roofline fraction on the level-L icosphere operator
(caller's vertex order and the solver-internal Morton-cell order), and the heat solve of config 4.

Usage (GPU box):  python tools/sweep_spmv.py [level]      -> gpurun_out/sweep_spmv.json
Round 1 used this script with environment-selected kernel variants (CTA shape, batching, cache
hints) to pick the shipped SpMV form; the losing variants were removed from the library.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, heat, mesh as M  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
_peaks = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(_peaks)).get("hbm_gbs", 6550.1) if os.path.exists(_peaks) else 6550.1
msh = M.icosphere(level)
ctx = _lib.default_context()
dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
a, b = _lib.assemble(ctx, dm, 0, False)
n, nnz = a.n, a.nnz
out = {"level": level, "n": n, "nnz": nnz, "peak_gb_s": PEAK}
for m, reps in ((1, 100), (2, 100), (64, 20)):
    nbytes = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m
    ms = _lib.spmm_benchmark(ctx, a, m, reps)
    msr = _lib.spmm_benchmark(ctx, a, m, reps, renumber=True)
    out[f"spmm_m{m}"] = {"ms": ms, "frac": nbytes / ms / 1e6 / PEAK, "renumbered_ms": msr, "renumbered_frac": nbytes / msr / 1e6 / PEAK}
    print(f"m={m}", out[f"spmm_m{m}"], flush=True)
for mult in (1.0, 16.0):
    heat.diffusion(msh, [0], m=mult)
    t0 = time.perf_counter()
    u = heat.diffusion(msh, [0], m=mult)
    out[f"heat_m{int(mult)}"] = {"wall_s": time.perf_counter() - t0, "info": heat.diffusion.last_info, "u0": float(u[0]), "usum": float(u.sum())}
    print(f"heat m={mult}", out[f"heat_m{int(mult)}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_spmv.json", "w"), indent=1)
