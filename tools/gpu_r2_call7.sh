#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/c7_pytest.log 2>&1
tail -70 gpurun_out/c7_pytest.log
timeout 300 python tools/spmm_shapes.py 9 121 > gpurun_out/c7_spmm_shapes.log 2>&1; grep -E "m=(8|16|24|32|48|64|128) " gpurun_out/c7_spmm_shapes.log | head -60
timeout 300 python tools/dense_bench.py > gpurun_out/c7_dense.log 2>&1; tail -30 gpurun_out/c7_dense.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c7_asm.log 2>&1; tail -4 gpurun_out/c7_asm.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; tail -3 gpurun_out/c7_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c7_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
        print(k, d.get(k))
    print("roofline", {k: v for k, v in d["roofline"].items() if k != "spmm_shapes_in_timed_region"})
    for s in d["roofline"]["spmm_shapes_in_timed_region"]: print("   ", s)
    print("classes", d["kernel_classes"])
    print("configs", json.dumps(d["configs"], indent=1))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/c7_bench.json").read()[:2000])
PY
