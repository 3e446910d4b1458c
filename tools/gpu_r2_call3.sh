#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
LAPY_B200_TRACE=1 timeout 120 python tools/asm_once.py ico9 3 2>&1 | grep -E "lb trace|assemble" | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tria_element|incidence_fill4|row_count|row_fill|scan_tile" -c 12 -o gpurun_out/asm_tria_r2 python tools/asm_once.py ico9 1 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
ncu -i gpurun_out/asm_tria_r2.ncu-rep --page raw --csv > gpurun_out/asm_tria_r2_raw.csv 2>/dev/null
ls -la gpurun_out/asm_tria_r2*
