#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/san_mem.log 2>&1; echo "memcheck rc $?"; grep -E "ERROR SUMMARY|Invalid|tri nnz|selftest|eigs|heat" gpurun_out/san_mem.log | head -12
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/san_race.log 2>&1; echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|Race reported|hazard" gpurun_out/san_race.log | sort | uniq -c | head -12
