#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cat > /tmp/mini.py <<'PY'
import numpy as np, lapy_b200, sys
from lapy_b200 import mesh as M, heat, diffgeo
m = M.icosphere(4)
fem = lapy_b200.Solver(m)
ev, evec = fem.eigs(k=10)
print("eigs", ev[:5], fem.last_info)
u = heat.diffusion(m, [0])
print("heat", u[:3], heat.diffusion.last_info)
g = diffgeo.compute_geodesic_f(m, u)
print("geo max", g.max())
fl = lapy_b200.Solver(m, lump=True)
x = fl.poisson(np.sin(m.v[:, 0]), dtup=(np.array([0, 5]), np.array([0.0, 1.0])))
print("poisson", x[:3], fl.last_info)
t = M.cube_tets(8)
ft = lapy_b200.Solver(t)
print("tet eigs", ft.eigs(k=6)[0], ft.last_info)
PY
LAPY_B200_TRACE=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/mini.py > gpurun_out/memcheck2.log 2>&1
echo "memcheck exit $?"; grep -v "^\[lb trace\]" gpurun_out/memcheck2.log | tail -30
LAPY_B200_TRACE=1 timeout 600 python /tmp/mini.py 2>&1 | grep -E "lobpcg it|AMG|n=|eigs|heat|geo|poisson|Error|error" | tail -60
timeout 1700 python -m pytest tests/test_solvers_gpu.py -m gpu -q 2>&1 | tail -60
