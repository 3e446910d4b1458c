#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_assembly_gpu.py tests/test_frows_gpu.py -m gpu -q -x ) > gpurun_out/k2_pytest.log 2>&1
tail -5 gpurun_out/k2_pytest.log
timeout 300 python tools/time_assembly.py 9 61 > gpurun_out/k2_asm.log 2>&1; tail -4 gpurun_out/k2_asm.log
timeout 300 python tools/spmm_once.py 64 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:"strip_rows|scan_tile" -c 8 -o gpurun_out/asm_strip_r2 python tools/asm_once.py ico9 1 > gpurun_out/k2_ncu.log 2>&1; tail -2 gpurun_out/k2_ncu.log
ncu -i gpurun_out/asm_strip_r2.ncu-rep --page raw --csv > gpurun_out/asm_strip_r2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/asm_strip_r2_raw.csv | grep -v scan_tile
