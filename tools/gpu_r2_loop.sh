#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for cfg in "2 3" "1 3" "1 4" "1 5"; do
set -- $cfg
echo "=== gamma $1 ncyc $2"
LAPY_B200_GAMMA=$1 LAPY_B200_NCYC=$2 LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py ico9 2 2>&1 | grep -E "phases|step 1" | tail -2 | cut -c1-250
LAPY_B200_GAMMA=$1 LAPY_B200_NCYC=$2 LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -E "phases|step 1" | tail -2 | cut -c1-250
done
