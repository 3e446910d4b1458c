#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 420 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/eigs_loop.py ico9 1 > gpurun_out/san_init3.log 2>&1; echo "rc $?"
grep -E "ERROR SUMMARY|step 0" gpurun_out/san_init3.log | cut -c1-200 | head
grep -E "Device Frame" gpurun_out/san_init3.log | sed 's/+0x[0-9a-f]* / /' | sed 's/(.*) in / in /' | sort | uniq -c | sort -rn | head -12
