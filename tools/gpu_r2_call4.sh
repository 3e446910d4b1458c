#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/c4_memcheck.log 2>&1 <<'PY'
import numpy as np, lapy_b200
from lapy_b200 import mesh as M
for mk in (lambda: M.icosphere(3), lambda: M.cube_tets(5)):
    m = mk()
    for lump in (False, True):
        f = lapy_b200.Solver(m, lump=lump)
        print(type(m).__name__, lump, f.stiffness.nnz, f.mass.nnz)
PY
echo "memcheck exit $?"; tail -8 gpurun_out/c4_memcheck.log
( time timeout 900 python -m pytest tests/test_assembly_gpu.py -m gpu -q -x ) > gpurun_out/c4_pytest.log 2>&1
tail -30 gpurun_out/c4_pytest.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c4_asm.log 2>&1; tail -4 gpurun_out/c4_asm.log
LAPY_B200_TRACE=1 timeout 120 python tools/asm_once.py ico9 2 2>&1 | grep -E "lb trace|assemble" | tail -9
