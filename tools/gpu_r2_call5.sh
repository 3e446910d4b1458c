#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tria_row|incidence_fill4|row_count_kernel|tria_element" -c 6 -o gpurun_out/asm_tria_r2b python tools/asm_once.py ico9 1 > gpurun_out/c5_ncu.log 2>&1
tail -3 gpurun_out/c5_ncu.log
ncu -i gpurun_out/asm_tria_r2b.ncu-rep --page raw --csv > gpurun_out/asm_tria_r2b_raw.csv 2>/dev/null
ncu -i gpurun_out/asm_tria_r2b.ncu-rep --page source --csv -k regex:tria_row_fill_fast > gpurun_out/asm_tria_r2b_fill_source.csv 2>/dev/null
ls -la gpurun_out/asm_tria_r2b*
