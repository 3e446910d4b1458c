"""Dev tool: A/B of the SpMM kernels (strip-staged fp64 / fp32, row-wise fallback) per block width.

    python tools/spmm_variants.py [level] [cube_n]   -> gpurun_out/spmm_variants.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib, mesh as M  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
cube_n = int(sys.argv[2]) if len(sys.argv) > 2 else 121
ctx = _lib.default_context()
out = {}
names = {0: "strip", 1: "rowwise", 2: "f32", 23: "f32-3cta", 24: "f32-4cta", 25: "f32-5cta"}
for name, msh, kind in (("ico%d" % level, M.icosphere(level), 0), ("cube%d" % cube_n, M.cube_tets(cube_n), 3)):
    dm = _lib.DeviceMesh(ctx, msh.v, msh.t)
    a, b = _lib.assemble(ctx, dm, kind, False)
    for m in (4, 8, 16, 28, 32, 40, 64, 78, 128):
        row = {}
        for var in (0, 1, 2, 23, 24, 25):
            ms = _lib.spmm_benchmark(ctx, a, m, 20, renumber=True, variant=var)
            es = 4 if var >= 2 else 8
            mm = (m + 3) // 4 * 4 if var >= 2 else m
            nbytes = (4.0 + es) * a.nnz + 4.0 * (a.n + 1) + 2.0 * es * a.n * mm
            row[names[var]] = {"ms": ms, "gb_s": nbytes / ms / 1e6}
        out[f"{name} m={m}"] = row
        print(name, m, "  ".join(f"{k} {v['ms']:.4f}ms/{v['gb_s']:.0f}GB/s" for k, v in row.items()), flush=True)
    del a, b, dm
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/spmm_variants.json", "w"), indent=1)
