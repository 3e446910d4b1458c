"""Golden outputs of the UNMODIFIED reference for the callers either side of the hot path (SURVEY.md
§8f.1-3): mesh pre-steps, tria_compute_divergence2, tria_compute_rotated_f, tria_mean_curvature_flow,
tria_spherical_project, heat / geodesics with an anisotropic operator.  Dev container only:

    python tools/make_golden_frows.py   ->  tests/golden/frows.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402

lapy, DATA = refshim.load()
from lapy import Solver, TetMesh, TriaMesh, diffgeo, heat  # noqa: E402

from lapy_b200 import mesh as M  # noqa: E402

d = {}
# 1. a bumpy ellipsoid (closed, oriented, axes y > z > x like a brain in FreeSurfer space)
s = M.perturbed_sphere(4, seed=3, amp=0.25)
v = s.v * np.array([1.0, 2.0, 1.5])
geo = TriaMesh(v, s.t)
d["ell_v"], d["ell_t"] = geo.v, geo.t
d["ell_avg_edge"] = np.float64(geo.avg_edge_length())
d["ell_area"], d["ell_volume"] = np.float64(geo.area()), np.float64(geo.volume())
vids, tids = geo.edges()
d["ell_edges_vids"], d["ell_edges_tids"] = vids, tids
d["ell_vertex_normals"] = geo.vertex_normals()
for k, arr in zip(("umin", "umax", "cmin", "cmax", "cmean", "cgauss", "normals"), geo.curvature(smoothit=3)):
    d["ell_curv_" + k] = arr
for k, arr in zip(("u1", "u2", "c1", "c2"), geo.curvature_tria(smoothit=10)):
    d["ell_curvtria_" + k] = arr
flow = diffgeo.tria_mean_curvature_flow(TriaMesh(geo.v, geo.t), max_iter=3)
d["ell_mcf3_v"] = flow.v
flow = diffgeo.tria_mean_curvature_flow(TriaMesh(geo.v, geo.t), max_iter=8, step=0.5)
d["ell_mcf8_step05_v"] = flow.v
sph = diffgeo.tria_spherical_project(TriaMesh(geo.v, geo.t), flow_iter=3)
d["ell_sphere_v"] = sph.v
rng = np.random.default_rng(0)
w = rng.normal(size=(3, 2))
f = np.sin(geo.v @ w)
d["ell_f"] = f
g1, g2 = diffgeo.tria_compute_gradient(geo, f[:, 0]), diffgeo.tria_compute_gradient(geo, f)
d["ell_div2_1d"] = diffgeo.tria_compute_divergence2(geo, g1)
d["ell_div2_2d"] = diffgeo.tria_compute_divergence2(geo, g2)
d["ell_rot_1d"] = diffgeo.tria_compute_rotated_f(geo, f[:, 0])
d["ell_rot_2d"] = diffgeo.tria_compute_rotated_f(geo, f)
# anisotropic operator through the public constructor (curvature_tria inside), heat with aniso
fem = Solver(geo, aniso=(1.0, 4.0), aniso_smooth=5)
d["ell_aniso_evals"] = fem.eigs(k=8)[0]
d["ell_aniso_Adata"], d["ell_aniso_Aindices"], d["ell_aniso_Aindptr"] = fem.stiffness.data, fem.stiffness.indices, fem.stiffness.indptr
d["ell_heat_aniso"] = heat.diffusion(geo, [0, 50], m=1.0, aniso=2.0)
# 1b. open mesh: the rotated-gradient function is non-trivial there (boundary flux)
sq = TriaMesh.read_off(DATA + "/square-mesh.off")
fs = np.column_stack((sq.v[:, 0].astype(np.float64), np.sin(3 * sq.v[:, 0]) * sq.v[:, 1]))
d["sq_f"] = fs
d["sq_rot_1d"] = diffgeo.tria_compute_rotated_f(sq, fs[:, 0])
d["sq_rot_2d"] = diffgeo.tria_compute_rotated_f(sq, fs)
d["sq_div2"] = diffgeo.tria_compute_divergence2(sq, diffgeo.tria_compute_gradient(sq, fs[:, 1]))
# 2. tet cube: adjacency + avg edge length
t = M.cube_tets(6)
tg = TetMesh(t.v, t.t)
a = tg.adj_sym.tocsc()
d["tet_adj_indptr"], d["tet_adj_indices"], d["tet_adj_data"] = a.indptr, a.indices, a.data
d["tet_avg_edge"] = np.float64(tg.avg_edge_length())
out = os.path.join(os.path.dirname(HERE), "tests", "golden", "frows.npz")
np.savez_compressed(out, **d)
print(out, os.path.getsize(out) / 1e6, "MB", len(d), "arrays")
