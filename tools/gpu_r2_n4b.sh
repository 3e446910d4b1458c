#!/bin/bash
# 4-GPU run of the driver's scaling command with the final code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "bench rc $?"; grep "\[bench\]" gpurun_out/n4_bench.err | cut -c1-120
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/n4_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"])
print(json.dumps(d["configs"])[:1200])
PY
