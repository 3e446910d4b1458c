#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc $?"; grep "\[bench\]" gpurun_out/k_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/k_bench.json"))
for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
    print(k, d.get(k))
print("roofline frac", d["roofline"]["frac"], "isolated", d["roofline"]["isolated_frac"])
print("configs", json.dumps(d["configs"])[:1500])
PY
