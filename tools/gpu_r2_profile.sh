#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/p_pytest.log 2>&1
tail -12 gpurun_out/p_pytest.log
timeout 300 python tools/dense_bench.py > gpurun_out/p_dense.log 2>&1; tail -34 gpurun_out/p_dense.log
timeout 300 python tools/time_assembly.py 9 121 > gpurun_out/p_asm.log 2>&1; tail -4 gpurun_out/p_asm.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc $?"; tail -5 gpurun_out/p_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/p_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
        print(k, d.get(k))
    print("roofline frac", d["roofline"]["frac"], d["roofline"]["kernel"][:80])
    print("classes", {k: round(v["ms_per_step"], 1) for k, v in d["kernel_classes"].items()})
    print("small dense", d["dense_shapes_in_timed_region"]["small_dense"][:4])
    print("configs", json.dumps(d["configs"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/p_bench.json").read()[:2000])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -c 6 -o gpurun_out/spmm_r2 python tools/spmm_once.py 64 > gpurun_out/p_ncu_spmm.log 2>&1; tail -2 gpurun_out/p_ncu_spmm.log
ncu -i gpurun_out/spmm_r2.ncu-rep --page raw --csv > gpurun_out/spmm_r2_raw.csv 2>/dev/null
python tools/ncu_to_json.py gpurun_out/spmm_r2_raw.csv 64 gpurun_out/ncu_dominant_kernel_r2.json
python tools/ncu_summary.py gpurun_out/spmm_r2_raw.csv | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 52000 --launch-count 17000 --csv --log-file gpurun_out/launches_shapedna_r2.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/p_launch_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_shapedna_r2.csv | head -25
