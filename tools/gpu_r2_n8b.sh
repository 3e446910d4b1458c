#!/bin/bash
# config 5 (batched ShapeDNA, level-7 surfaces) on 8 GPUs: workers per GPU 2 / 3 / 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc
for w in 3 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$w tools/bench_batch.py --meshes 384 --workers $w 2>/dev/null | tail -1 | tee -a gpurun_out/n8_batch.log
done
