#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/dense_bench.py 2>&1 | tail -28
timeout 300 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x -k "dmma" 2>&1 | tail -2
