#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_frows_gpu.py -m gpu -q -x ) 2>&1 | tail -5
timeout 300 python tools/time_assembly.py 9 61 2>&1 | tail -4 | cut -c1-200
