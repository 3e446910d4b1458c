#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for nc in 1 2 3; do
echo "=== NCYC $nc"
LAPY_B200_NCYC=$nc LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico9 2 2>&1 | grep -E "phases|step|nested" | tail -2
LAPY_B200_NCYC=$nc LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -E "phases|step|nested" | tail -2
LAPY_B200_NCYC=$nc LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico7 3 2>&1 | grep -E "phases|step|nested" | tail -2
done
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -2
