#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
LAPY_B200_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "rc $?"
grep -n "per-call e2e\|per-step wall" gpurun_out/e_bench.err
# phases of the last four eigs calls before the e2e report (3 timed e2e calls + last warm-up)
grep -n "enter assemble\|strip rows\|AMG setup\|coarse mass\|nested coarse\|fine-level\|un-renumber\|enter eigs\|work blocks" gpurun_out/e_bench.err | tail -44 | cut -c1-90
