"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

    python tools/make_golden.py

Every array stored here is an input handed to, or an output produced by, the reference's own
functions (lapy.Solver, lapy.heat.diffusion, lapy.diffgeo.*, lapy.shapedna.compute_shapedna,
TriaMesh.curvature_tria, *.avg_edge_length) imported from a scratch copy of /root/reference
(tools/refshim.py).  The oracle (oracle/) and the CUDA path are both tested against these files;
nothing at test/bench time reads /root/reference.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import refshim  # noqa: E402

lapy, DATA = refshim.load()
from lapy import Solver, TetMesh, TriaMesh, diffgeo, heat, shapedna  # noqa: E402

from lapy_b200 import mesh as M  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def csc(d, name, m):
    m = m.tocsc()
    assert m.has_canonical_format or True
    d[name + "_indptr"] = m.indptr
    d[name + "_indices"] = m.indices
    d[name + "_data"] = m.data
    d[name + "_shape"] = np.array(m.shape)


def read_ev(path):
    txt = open(path).read()
    blk = txt.split("Eigenvalues:")[1].split("Eigenvectors:")[0]
    vals = blk.replace("{", " ").replace("}", " ").replace(";", " ").split()
    return np.array([float(x) for x in vals])


def common(d, geo, k, store_evecs=True, lump_eigs=False):
    d["v"], d["t"] = geo.v, geo.t
    d["avg_edge_length"] = np.float64(geo.avg_edge_length())
    for lump in (False, True):
        fem = Solver(geo, lump=lump)
        tag = "lump" if lump else "full"
        if not lump:
            csc(d, "A", fem.stiffness)
        csc(d, "B_" + tag, fem.mass)
        if lump == lump_eigs:
            ev, evec = fem.eigs(k=k)
            d["evals"] = ev
            d["evals_lump"] = np.array(lump)
            if store_evecs:
                d["evecs"] = evec


def field(geo, seed=0):
    rng = np.random.default_rng(seed)
    w = rng.normal(size=(3, 3))
    return np.sin(geo.v.astype(np.float64) @ w).astype(np.float64)  # (n, 3) smooth functions


def geodesics(d, geo, seeds):
    u = heat.diffusion(geo, seeds, m=1.0)
    d["heat_seeds"] = np.asarray(seeds)
    d["heat_u"] = u
    d["geodesic"] = diffgeo.compute_geodesic_f(geo, u)
    f = field(geo)
    d["f"] = f
    d["grad_1d"] = diffgeo.compute_gradient(geo, f[:, 0])
    d["grad_2d"] = diffgeo.compute_gradient(geo, f)
    d["div_1d"] = diffgeo.compute_divergence(geo, d["grad_1d"])
    d["div_2d"] = diffgeo.compute_divergence(geo, d["grad_2d"])
    d["geodesic_2d"] = diffgeo.compute_geodesic_f(geo, np.column_stack((u, f[:, 0])))


def poisson_cases(d, geo, lump=True):
    fem = Solver(geo, lump=lump)
    f = field(geo, 1)
    f = f - (fem.mass @ f).sum(0) / fem.mass.sum()  # compatible rhs on closed meshes
    d["poisson_h"] = f
    d["poisson_lump"] = np.array(lump)
    d["poisson_1d"] = fem.poisson(f[:, 0])
    d["poisson_2d"] = fem.poisson(f)
    didx = np.array([0, 1, 5])
    dval = np.array([0.0, 0.5, -0.25])
    d["poisson_didx"], d["poisson_dval"] = didx, dval
    d["poisson_dirichlet_1d"] = fem.poisson(f[:, 0], dtup=(didx, dval))
    d["poisson_dirichlet_2d"] = fem.poisson(f, dtup=(didx, dval))
    d["poisson_laplace_dirichlet"] = fem.poisson(0.0, dtup=(didx, dval))
    nidx = np.array([2, 3])
    nval = np.array([1.0, -1.0])
    d["poisson_nidx"], d["poisson_nval"] = nidx, nval
    d["poisson_neumann_dirichlet"] = fem.poisson(f[:, 0], dtup=(didx, dval), ntup=(nidx, nval))


def aniso_case(d, geo, aniso=(2.0, 5.0), smooth=3):
    u1, u2, c1, c2 = geo.curvature_tria(smoothit=smooth)
    am = np.empty((geo.t.shape[0], 2))
    am[:, 1] = np.exp(-aniso[1] * np.abs(c1))
    am[:, 0] = np.exp(-aniso[0] * np.abs(c2))
    d["aniso_u1"], d["aniso_u2"], d["aniso_mat"] = u1, u2, am
    d["aniso"], d["aniso_smooth"] = np.array(aniso), np.array(smooth)
    a, b = Solver._fem_tria_aniso(geo, u1, u2, am, lump=False)
    csc(d, "A_aniso", a)
    csc(d, "B_aniso", b)
    fem = Solver(geo, aniso=aniso, aniso_smooth=smooth)
    assert (fem.stiffness != a).nnz == 0
    d["aniso_evals"] = fem.eigs(k=6)[0]


def save(name, d):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(d)} arrays")


def main():
    # 1. cubeTria.vtk (float32 vertices) - BASELINE.json config 1
    geo = TriaMesh.read_vtk(DATA + "/cubeTria.vtk")
    d = {}
    common(d, geo, k=10)
    d["ev_file"] = read_ev(DATA + "/cubeTria.ev")
    for lump in (False, True):
        csc(d, "M_" + ("lump" if lump else "full"), Solver.fem_tria_mass(geo, lump=lump))
    geodesics(d, geo, [0])
    poisson_cases(d, geo, lump=False)
    save("cubeTria", d)

    # 2. square-mesh.off (open mesh, float32) - the reference's solver/heat/geodesic tests
    geo = TriaMesh.read_off(DATA + "/square-mesh.off")
    d = {}
    common(d, geo, k=10, lump_eigs=True)
    bnd = np.concatenate(geo.boundary_loops())
    d["boundary"] = bnd
    geodesics(d, geo, bnd)
    d["heat_multi_seeds_1"], d["heat_multi_seeds_2"] = np.array([0]), np.array([1, 2])
    d["heat_multi"] = heat.diffusion(geo, [bnd, np.array([0]), np.array([1, 2])], m=1)
    poisson_cases(d, geo, lump=True)
    save("squareMesh", d)

    # 3. cubeTetra.vtk (float32, 24,000 negatively oriented tets)
    geo = TetMesh.read_vtk(DATA + "/cubeTetra.vtk")
    d = {}
    common(d, geo, k=10, store_evecs=False)
    d["ev_file"] = read_ev(DATA + "/cubeTetra.ev")
    geodesics(d, geo, [0])
    save("cubeTetra", d)

    # 4. icosphere level 3 (float64) incl. anisotropic operator
    s = M.icosphere(3)
    geo = TriaMesh(s.v, s.t)
    d = {}
    common(d, geo, k=20)
    for lump in (False, True):
        csc(d, "M_" + ("lump" if lump else "full"), Solver.fem_tria_mass(geo, lump=lump))
    geodesics(d, geo, [0, 7])
    poisson_cases(d, geo, lump=True)
    aniso_case(d, geo)
    save("ico3", d)

    # 5. torus.off (float32, genus 1, valence-regular) with anisotropy
    geo = TriaMesh.read_off(DATA + "/torus.off")
    d = {}
    common(d, geo, k=10)
    aniso_case(d, geo, aniso=(1.0, 10.0), smooth=2)
    save("torus", d)

    # 6. icosphere level 5 (10,242 v): k=50 spectrum (cuts the l=7 cluster) + CSR
    s = M.icosphere(5)
    geo = TriaMesh(s.v, s.t)
    d = {}
    common(d, geo, k=50, store_evecs=False)
    sd = shapedna.compute_shapedna(geo, k=50)
    d["shapedna_evals"] = sd["Eigenvalues"]
    d["shapedna_meta"] = np.array([sd["Refine"], sd["Degree"], sd["Dimension"], sd["Elements"], sd["DoF"], sd["NumEW"]])
    geodesics(d, geo, [0])
    save("ico5", d)

    # 7. structured tet cube n=9 (float64)
    s = M.cube_tets(9)
    geo = TetMesh(s.v, s.t)
    d = {}
    common(d, geo, k=12)
    geodesics(d, geo, [0])
    poisson_cases(d, geo, lump=True)
    save("cube9", d)

    # 8. degenerate / ragged triangle soup (float64): zero-area triangle, repeated vertex,
    #    unreferenced trailing vertex (shape is inferred from max index),
    #    non-manifold edge.  Assembly only.
    v = np.array(
        [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5], [2, 0, 0], [0.5, 0.5, -1], [3, 3, 3], [9, 9, 9]],
        dtype=np.float64,
    )
    t = np.array([[0, 1, 2], [1, 3, 2], [0, 1, 4], [0, 1, 5], [1, 2, 5], [2, 3, 6], [1, 4, 3]])
    geo = TriaMesh(v, t)
    d = {"v": v, "t": t}
    for lump in (False, True):
        a, b = Solver._fem_tria(geo, lump=lump)
        csc(d, "A", a)
        csc(d, "B_" + ("lump" if lump else "full"), b)
        csc(d, "M_" + ("lump" if lump else "full"), Solver.fem_tria_mass(geo, lump=lump))
    save("degenerate", d)

    # 9. level-7 / level-8 / level-9 k=50 spectra quoted in BASELINE.md §5.2 are stored by
    #    tools/make_golden_spectra.py (text -> npz), not recomputed here (26 min at level 9).


if __name__ == "__main__":
    main()
