#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for ct in 1e-3 1e-2 5e-2; do
echo "=== coarse tol $ct"
LAPY_B200_CTOL=$ct LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico9 2 2>&1 | grep -E "phases|step 1|nested|second Gram" | tail -7 | cut -c1-250
done
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -E "phases|step 1|nested|second Gram" | tail -5 | cut -c1-250
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico7 3 2>&1 | grep -E "phases|step 2|second Gram" | tail -3 | cut -c1-250
timeout 600 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -2
