"""Debug driver for the row-partitioned mode: torchrun --nproc-per-node 2 tools/rowpart_dbg.py [level]

Variants by environment: LAPY_B200_HALO=1 (boundary-only halo exchange), LAPY_B200_DIST_AMG=full (replicated
hierarchy applied column-parallel + nested start)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(240, exit=True)
os.environ.setdefault("LAPY_B200_TRACE", "1")
os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print(f"[rank {rank} {time.strftime('%X')}]", *a, file=sys.stderr, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg up")
import lapy_b200
from lapy_b200 import _lib, mesh as M
ctx = _lib.Context(local)
ctx.init_row_partition()
log("comm up")
what = sys.argv[1] if len(sys.argv) > 1 else "6"
lvl = int(what) if what.isdigit() else -1
mesh = M.icosphere(lvl) if lvl >= 0 else M.cube_tets(int(what[4:]))
fem = lapy_b200.Solver(mesh, ctx=ctx)
log("assembled")
errs = np.zeros(2)
_lib.check(_lib.lib().lb_dist_selftest(ctx.handle, fem._device("a").handle, _lib.ptr(errs)))
log("selftest spmm/gram max abs diff", errs)
t0 = time.perf_counter()
try:
    ev, evec = fem.eigs(k=50)
except Exception as e:
    log('eigs failed', e); ev = np.zeros(50)
log("eigs done", time.perf_counter() - t0, fem.last_info, ev[:4])
t0 = time.perf_counter()
try:
    ev, evec = fem.eigs(k=50)
    log("second eigs", time.perf_counter() - t0, fem.last_info, "lam1/pi^2", ev[1] / np.pi**2)
except Exception as e:
    log('second eigs failed', e)

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "spectra.npz"))
key = f"ico{lvl}_k50"
if key in g:
    log("max rel err", np.max(np.abs(ev[1:] - g[key][1:]) / g[key][1:]))
ctx.leave_row_partition()
dist.destroy_process_group()
