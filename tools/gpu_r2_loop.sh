#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -2
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico9 3 2>&1 | grep -E "phases|step|nested" | tail -5 | cut -c1-250
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -E "phases|step 1" | tail -2 | cut -c1-250
python tools/trace_eigs.py ico9 2>&1 | grep -E "max rel err"
