#!/bin/bash
# 4-GPU validation of the row-partitioned eigensolve (halo exchange + column-parallel single-precision cycle)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_rowpart_gpu.py -m gpu -q -x > gpurun_out/n4_pytest.log 2>&1; tail -3 gpurun_out/n4_pytest.log
for np in 2 4; do
for what in 9 cube121; do
    echo "=== $what on $np GPUs"
    LAPY_B200_TRACE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 2951$np tools/rowpart_dbg.py $what > gpurun_out/rp${np}_${what}.log 2>&1
    grep -E "eigs done|second eigs|max rel|failed|Error|error|selftest" gpurun_out/rp${np}_${what}.log | grep "rank 0" | cut -c1-330
done
done
