#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
LAPY_B200_TRACE=1 timeout 300 python tools/eigs_loop.py cube121 2 2>&1 | grep -v "lobpcg it" | tail -24
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc $?"; tail -5 gpurun_out/p_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/p_bench.json"))
    for k in ("value", "ms_per_step", "parity", "assembly", "e2e", "eigs", "gpu_launches"):
        print(k, d.get(k))
    print("roofline frac", d["roofline"]["frac"], d["roofline"]["kernel"][:120], "isolated", d["roofline"]["isolated_frac"])
    print("classes", {k: round(v["ms_per_step"], 1) for k, v in d["kernel_classes"].items()}, d["kernel_classes_note"])
    print("small dense", d["dense_shapes_in_timed_region"]["small_dense"][:4])
    print("configs", json.dumps(d["configs"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/p_bench.json").read()[:2000])
PY
for v in 2 0; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_strip_kernel -c 6 -o gpurun_out/spmm_r2_v$v -f python tools/spmm_once.py 64 $v > gpurun_out/p_ncu_spmm_v$v.log 2>&1; tail -2 gpurun_out/p_ncu_spmm_v$v.log
ncu -i gpurun_out/spmm_r2_v$v.ncu-rep --page raw --csv > gpurun_out/spmm_r2_v${v}_raw.csv 2>/dev/null
python tools/ncu_to_json.py gpurun_out/spmm_r2_v${v}_raw.csv 64 gpurun_out/ncu_spmm_strip_v$v.json
python tools/ncu_summary.py gpurun_out/spmm_r2_v${v}_raw.csv | tail -1
done
