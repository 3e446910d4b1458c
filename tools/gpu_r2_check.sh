#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/k_pytest.log 2>&1
tail -12 gpurun_out/k_pytest.log
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico9 2 2>&1 | grep -E "phases|step|nested" | tail -4
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -E "phases|step|nested" | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc $?"; tail -5 gpurun_out/k_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
        print(k, d.get(k))
    print("roofline frac", d["roofline"]["frac"], d["roofline"]["kernel"][:80], "isolated", d["roofline"]["isolated_frac"])
    print("classes", {k: round(v["ms_per_step"], 1) for k, v in d["kernel_classes"].items()}, d["kernel_classes_note"])
    print("configs", json.dumps(d["configs"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/k_bench.json").read()[:2000])
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
