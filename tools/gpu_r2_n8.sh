#!/bin/bash
# 8-GPU validation of the driver's scaling run: mesh-parallel bench + rowpart + batch records
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench rc $?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/n8_bench.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/n8_bench.json") if l.startswith("{")][-1])
    for k in ("value", "ms_per_step", "n_gpus", "parity", "e2e", "eigs"):
        print(k, d.get(k))
    print("configs", json.dumps(d["configs"], indent=1))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/n8_bench.json").read()[:3000])
PY
