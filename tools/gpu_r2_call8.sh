#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_solvers_gpu.py -m gpu -q -x ) > gpurun_out/c8_pytest.log 2>&1
tail -6 gpurun_out/c8_pytest.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c8_asm.log 2>&1; tail -4 gpurun_out/c8_asm.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "bench rc $?"; tail -30 gpurun_out/c8_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c8_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
        print(k, d.get(k))
    print("roofline", {k: v for k, v in d["roofline"].items() if k != "spmm_shapes_in_timed_region"})
    for s in d["roofline"]["spmm_shapes_in_timed_region"]: print("   ", s)
    print("classes", d["kernel_classes"])
    print("dense", json.dumps(d["dense_shapes_in_timed_region"]))
    print("configs", json.dumps(d["configs"], indent=1))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/c8_bench.json").read()[:2000])
PY
for w in 4 8; do timeout 300 python tools/bench_batch.py --meshes 32 --workers $w 2>&1 | tail -1; done
