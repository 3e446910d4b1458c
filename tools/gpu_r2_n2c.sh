#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/n2c_bench.json 2> gpurun_out/n2c_bench.err; echo "bench rc $?"; grep "\[bench\]" gpurun_out/n2c_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/n2c_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"])
PY
