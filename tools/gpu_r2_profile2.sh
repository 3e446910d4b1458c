#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x -k "lumped_and_small" 2>&1 | tail -2
timeout 300 python tools/bench_batch.py 2>&1 | tail -6
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 ) > gpurun_out/q_ref.json 2> gpurun_out/q_ref.err; tail -3 gpurun_out/q_ref.err; cut -c1-900 gpurun_out/q_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shapedna_r2.csv python tools/one_step.py 9 > gpurun_out/q_launch.log 2>&1; tail -2 gpurun_out/q_launch.log
python tools/launch_summary.py gpurun_out/launches_shapedna_r2.csv --after-last "strip_rows_kernel<0" | head -34
