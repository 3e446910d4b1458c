#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
what=${1:-ico8}
run() { echo "=== $*"; env "$@" python tools/trace_eigs.py $what 2>&1 | grep -E "phases|max rel|second eigs|AMG:|  n=" | tail -9; }
run LAPY_B200_CHEB=2
run LAPY_B200_CHEB=3
run LAPY_B200_CHEB=4
run LAPY_B200_CHEB=2 LAPY_B200_VCYCLES=2
run LAPY_B200_CHEB=2 LAPY_B200_BLOCK=80
run LAPY_B200_CHEB=2 LAPY_B200_TOL=1e-7
