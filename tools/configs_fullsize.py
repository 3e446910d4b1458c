"""BASELINE.json configs 3 and 4 at full size on one GPU: timings + size-independent checks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lapy_b200
from lapy_b200 import mesh as M, heat, diffgeo, _lib

out = {}
ctx = _lib.default_context()
def timed(f):
    ctx.sync(); t0 = time.perf_counter(); r = f(); ctx.sync(); return r, time.perf_counter() - t0

# config 3: cube121 tets
mesh, t = timed(lambda: M.cube_tets(121))
fem, t_asm = timed(lambda: lapy_b200.Solver(mesh))
fem2, t_asm2 = timed(lambda: lapy_b200.Solver(mesh))
dm = mesh._lb_device_mesh[2]
ms = []
for _ in range(5):
    dm.drop_cache(); ctx.timer_start(); a, b = _lib.assemble(ctx, dm, _lib.FEM_TETRA, False); ms.append(ctx.timer_stop())
nt, nv, nnz = mesh.t.shape[0], mesh.v.shape[0], a.nnz
algo = 16 * nt + 24 * nv + 2 * (12 * nnz + 4 * (nv + 1))
(ev, evec), t_eig = timed(lambda: fem.eigs(k=50))
(ev, evec), t_eig2 = timed(lambda: fem.eigs(k=50))
pi2 = np.pi ** 2
out["cube121"] = dict(nv=nv, nt=nt, nnz=int(nnz), assembly_device_ms=float(np.median(ms)), gelem_per_s=nt / np.median(ms) / 1e6,
                      roofline_frac=algo / (np.median(ms) * 1e-3) / 1e9 / 6550.1, solver_wall_s=t_asm2, eigs_s=t_eig2, info=fem.last_info,
                      ev_first=ev[:5].tolist(), ev1_over_pi2=ev[1] / pi2)
print(json.dumps(out["cube121"]), flush=True)
del fem, fem2, a, b, evec
# config 4: heat + geodesic on the level-9 icosphere
mesh = M.icosphere(9)
(u), t_heat = timed(lambda: heat.diffusion(mesh, [0], m=1.0))
(u), t_heat2 = timed(lambda: heat.diffusion(mesh, [0], m=1.0))
hinfo = heat.diffusion.last_info
# m=1 (t = h^2 = 5.6e-6): u underflows beyond ~1.7 rad and (grad u)^2 beyond ~0.8 rad, for the
# reference as well (np.sqrt((gradf**2).sum(1)) -> 0 -> inf -> nan_to_num): the heat method needs a
# larger t at this resolution.  m=16 gives sqrt(t) = 4h, the diffusion length of level 7 with m=1.
(u16), t_heat16 = timed(lambda: heat.diffusion(mesh, [0], m=16.0))
hinfo16 = heat.diffusion.last_info
(g), t_geo = timed(lambda: diffgeo.compute_geodesic_f(mesh, u16))
(g), t_geo2 = timed(lambda: diffgeo.compute_geodesic_f(mesh, u16))
far = int(np.argmin(mesh.v @ mesh.v[0]))
exact = np.arccos(np.clip(mesh.v @ mesh.v[0], -1, 1))
out["heat_geodesic_L9"] = dict(heat_m1_s=t_heat2, heat_m1_info=hinfo, u0=float(u[0]), usum=float(u.sum()), u_min=float(u.min()),
                               heat_m16_s=t_heat16, heat_m16_info=hinfo16, geodesic_s=t_geo2, geodesic_max=float(g.max()),
                               pi=float(np.pi), geodesic_at_antipode=float(g[far]),
                               max_abs_err_vs_great_circle=float(np.abs(g - exact).max()))
print(json.dumps(out["heat_geodesic_L9"]), flush=True)
json.dump(out, open("gpurun_out/configs_fullsize_r1.json", "w"), indent=1)
