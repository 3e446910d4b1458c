#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 1 2; do CUDA_LAUNCH_BLOCKING=1 LAPY_B200_TRACE=1 timeout 300 python tools/heat_once.py 9 2>&1 | grep -E "heat m=|geodesic|Error|error" | cut -c1-200; done
timeout 300 python tools/heat_once.py 9 2>&1 | grep -E "heat m=|geodesic|Error|error" | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/heat_once.py 8 2>&1 | grep -E "ERROR SUMMARY|Invalid" | head -3
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x -k "heat or geodesic or poisson or diffusion" 2>&1 | tail -2
