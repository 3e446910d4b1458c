"""Dev tool (round 2): A/B the opt-in wide-SpMM variants and vertex orders in ONE process.

    python tools/sweep_spmm_variants.py [level] [--eigs]   -> gpurun_out/sweep_spmm_variants.json

For every order in {default 128^3 bins, LAPY_B200_ORDER=fine} and every LAPY_B200_SPMM variant:
bit-check y = A x (64 columns, level-7 operator) against the default kernel, then the device time
of the 64-column product on the level-`level` operator in the solver numbering; with --eigs also a
full ShapeDNA k=50 (iterations, seconds).
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lapy_b200  # noqa: E402
from lapy_b200 import _lib, mesh as M  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 9
do_eigs = "--eigs" in sys.argv
VARIANTS = ["", "grouped2", "grouped", "grouped8", "grouped2s", "groupeds", "grouped8s"]
PEAK = 6550.1
ctx = _lib.default_context()
big, small = M.icosphere(level), M.icosphere(7)
rng = np.random.default_rng(0)
xs = rng.standard_normal((len(small.v), 64))
out = {}
for order in ("", "fine"):
    os.environ["LAPY_B200_ORDER"] = order
    os.environ["LAPY_B200_SPMM"] = ""
    dm_s = _lib.DeviceMesh(ctx, small.v, small.t)  # a fresh mesh -> a fresh ordering under this setting
    a_s, _ = _lib.assemble(ctx, dm_s, 0, False)
    dm = _lib.DeviceMesh(ctx, big.v, big.t)
    a, b = _lib.assemble(ctx, dm, 0, False)
    nbytes = 12.0 * a.nnz + 4.0 * (a.n + 1) + 16.0 * a.n * 64
    ref = None
    for var in VARIANTS:
        os.environ["LAPY_B200_SPMM"] = var
        y = _lib.spmm(ctx, a_s, xs)
        if ref is None:
            ref = y
        rec = {"bit_identical": bool(np.array_equal(y, ref)), "max_abs_diff": float(np.abs(y - ref).max())}
        ms = _lib.spmm_benchmark(ctx, a, 64, 20, renumber=True)
        rec.update(ms=ms, roofline_frac=nbytes / ms / 1e6 / PEAK)
        if do_eigs:
            t0 = time.perf_counter()
            ev, _, info = _lib.eigs(ctx, a, b, 50, -0.01)
            rec.update(eigs_s=time.perf_counter() - t0, iterations=info["iterations"], ev1=float(ev[1]))
        out[f"order={order or 'bins'} spmm={var or 'csr'}"] = rec
        print(f"order={order or 'bins':5s} spmm={var or 'csr':10s}", rec, flush=True)
    del a, b, a_s
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_spmm_variants.json", "w"), indent=1)
