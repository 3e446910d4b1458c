#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/eigs_loop.py ico9 3 2>&1 | tail -2 | cut -c1-260
timeout 300 python tools/bench_batch.py --meshes 48 --workers 4 2>&1 | tail -1
