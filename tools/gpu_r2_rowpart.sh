#!/bin/bash
# 2-GPU validation of the row-partitioned eigensolve: all three communication variants on the level-6
# test, then the level-9 icosphere and the 121^3 tet cube with the halo exchange + replicated hierarchy
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
LAPY_B200_TEST_HALO=1 timeout 600 python -m pytest tests/test_rowpart_gpu.py -m gpu -q -x > gpurun_out/rp_pytest.log 2>&1; tail -15 gpurun_out/rp_pytest.log
for what in 9 cube121; do
  for env in "LAPY_B200_HALO=0" "LAPY_B200_HALO=1" "LAPY_B200_HALO=1 LAPY_B200_DIST_AMG=full"; do
    echo "=== $what $env"
    env $env LAPY_B200_TRACE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/rowpart_dbg.py $what > gpurun_out/rp_${what}_$(echo $env | tr ' =' '__').log 2>&1
    grep -E "eigs done|second eigs|max rel|failed|Error|error" gpurun_out/rp_${what}_$(echo $env | tr ' =' '__').log | grep "rank 0" | cut -c1-400
  done
done
