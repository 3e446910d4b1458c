#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
what=${1:-ico8}
run() { echo "=== $*"; env "$@" python tools/trace_eigs.py $what 2>&1 | grep -E "phases|max rel|second eigs|nested level" | tail -7; }
run LAPY_B200_X=1
run LAPY_B200_NONESTED=1
