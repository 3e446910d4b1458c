#!/bin/bash
# two concurrent single-GPU processes: e2e loop with and without the page-locked result pool
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for p in 1 0; do
  echo "== POOL=$p, two processes"
  (LAPY_B200_DEVICE=0 POOL=$p timeout 300 python tools/e2e_loop.py > gpurun_out/e2e_p${p}_g0.log 2>&1 &
   LAPY_B200_DEVICE=1 POOL=$p timeout 300 python tools/e2e_loop.py > gpurun_out/e2e_p${p}_g1.log 2>&1 &
   wait)
  for g in 0 1; do tail -4 gpurun_out/e2e_p${p}_g$g.log | tr '\n' ' '; echo; done
done
