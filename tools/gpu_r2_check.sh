#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/k_pytest.log 2>&1
tail -12 gpurun_out/k_pytest.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/k_asm.log 2>&1; tail -4 gpurun_out/k_asm.log | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc $?"; tail -5 gpurun_out/k_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "spmv", "e2e", "eigs", "gpu_launches", "clocks"):
        print(k, d.get(k))
    print("roofline frac", d["roofline"]["frac"], d["roofline"]["kernel"][:80], "isolated", d["roofline"]["isolated_frac"], "traffic", d["roofline"]["traffic"])
    print("classes", {k: round(v["ms_per_step"], 1) for k, v in d["kernel_classes"].items()}, d["kernel_classes_note"])
    print("configs", json.dumps(d["configs"]))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/k_bench.json").read()[:2000])
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
