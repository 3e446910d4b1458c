#!/bin/bash
# round 2, first GPU call: the full GPU suite incl. the new full-size parity tests and the opt-in
# kernel variants written at the end of round 1, then the A/B measurements that decide which of
# them become the default.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
( time LAPY_B200_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/c1_pytest.log 2>&1
tail -40 gpurun_out/c1_pytest.log
LAPY_B200_TET_ROWS=fused LAPY_B200_TRIA_ROWS=fused timeout 600 python -m pytest tests/test_assembly_gpu.py -m gpu -q -x > gpurun_out/c1_pytest_fused.log 2>&1
tail -5 gpurun_out/c1_pytest_fused.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c1_asm_default.log 2>&1; tail -4 gpurun_out/c1_asm_default.log
LAPY_B200_TET_ROWS=fused LAPY_B200_TRIA_ROWS=fused timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/c1_asm_fused.log 2>&1; tail -4 gpurun_out/c1_asm_fused.log
timeout 300 python tools/spmm_shapes.py 9 121 > gpurun_out/c1_spmm_shapes.log 2>&1; tail -50 gpurun_out/c1_spmm_shapes.log
timeout 600 python tools/sweep_spmm_variants.py 9 --eigs > gpurun_out/c1_sweep_variants.log 2>&1; tail -16 gpurun_out/c1_sweep_variants.log
LAPY_B200_TRACE=1 timeout 300 python tools/trace_eigs.py ico9 > gpurun_out/c1_trace_ico9.log 2>&1; grep -E "phases|max rel|eigs|nested|AMG|n=" gpurun_out/c1_trace_ico9.log | tail -30
LAPY_B200_EIG=syevj LAPY_B200_TRACE=1 timeout 300 python tools/trace_eigs.py ico9 > gpurun_out/c1_trace_ico9_syevj.log 2>&1; grep -E "phases|max rel|second eigs" gpurun_out/c1_trace_ico9_syevj.log | tail -6
LAPY_B200_TRACE=1 timeout 300 python tools/trace_eigs.py ico7 > gpurun_out/c1_trace_ico7.log 2>&1; grep -E "phases|max rel|eigs|nested" gpurun_out/c1_trace_ico7.log | tail -10
