#!/bin/bash
# first-contact run on the B200 box: memcheck a small assembly, then the GPU test-suite
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/memcheck.log 2>&1 <<'PY'
import numpy as np, lapy_b200
from lapy_b200 import mesh as M
for mk in (lambda: M.icosphere(3), lambda: M.cube_tets(5)):
    m = mk()
    for lump in (False, True):
        f = lapy_b200.Solver(m, lump=lump)
        print(type(m).__name__, lump, f.stiffness.nnz, f.mass.nnz)
PY
echo "memcheck exit $?"; tail -15 gpurun_out/memcheck.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
