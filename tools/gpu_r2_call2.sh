#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/c2_memcheck.log 2>&1 <<'PY'
import numpy as np, lapy_b200
from lapy_b200 import mesh as M
for mk in (lambda: M.icosphere(3), lambda: M.cube_tets(5)):
    m = mk()
    for lump in (False, True):
        f = lapy_b200.Solver(m, lump=lump)
        print(type(m).__name__, lump, f.stiffness.nnz, f.mass.nnz)
f = lapy_b200.Solver(M.icosphere(4))
print(f.eigs(k=5)[0])
PY
echo "memcheck exit $?"; tail -12 gpurun_out/c2_memcheck.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/c2_pytest.log 2>&1
tail -60 gpurun_out/c2_pytest.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c2_asm.log 2>&1; tail -4 gpurun_out/c2_asm.log
LAPY_B200_TRACE=1 timeout 120 python tools/time_assembly.py 9 121 2>&1 | grep "lb trace" | tail -40 > gpurun_out/c2_asm_trace.log; tail -22 gpurun_out/c2_asm_trace.log
