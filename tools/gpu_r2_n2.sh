#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rowpart_gpu.py tests/test_solvers_gpu.py -m gpu -q -x -k "row_partitioned or dmma" > gpurun_out/n2_pytest.log 2>&1; tail -4 gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc $?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/n2_bench.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/n2_bench.json") if l.startswith("{")][-1])
    for k in ("value", "ms_per_step", "n_gpus", "parity", "e2e", "eigs"):
        print(k, d.get(k))
    print("configs", json.dumps(d["configs"], indent=1))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/n2_bench.json").read()[:3000])
PY
