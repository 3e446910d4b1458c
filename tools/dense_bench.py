"""Dev tool: device time of the fp64 DMMA block products at the shapes LOBPCG uses, the register-only
DMMA peak probe, and the update-kernel variants (k-chunks per barrier).  -> gpurun_out/dense_bench.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import _lib
ctx = _lib.default_context()
n = 2621442
out = {"dmma_peak_tflops": [_lib.dense_benchmark(ctx, 1, 1, 1, 2, 0, 5) for _ in range(3)]}
print("DMMA register-only peak TFLOP/s:", out["dmma_peak_tflops"], flush=True)
for p, q in ((192, 128), (128, 64), (64, 64), (94, 78), (128, 32), (128, 24), (128, 16), (128, 8), (64, 24)):
    for var, name in ((0, "default"), (2, "cp.async-ring")):
        if var == 2 and q <= 32:
            continue
        ms = _lib.dense_benchmark(ctx, n, p, q, 1, var, 10)
        tf = 2.0 * n * p * q / ms / 1e9
        out[f"update p={p} q={q} {name}"] = {"ms": ms, "tflops": tf}
        print(f"update p={p:3d} q={q:3d} {name}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
for p, q in ((192, 192), (128, 64), (64, 64), (94, 94), (128, 32), (128, 16), (79, 16), (16, 16)):
    ms = _lib.dense_benchmark(ctx, n, p, q, 0, 0, 10)
    tf = 2.0 * n * p * q / ms / 1e9
    out[f"gram p={p} q={q}"] = {"ms": ms, "tflops": tf}
    print(f"gram   p={p:3d} q={q:3d}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dense_bench.json", "w"), indent=1)
