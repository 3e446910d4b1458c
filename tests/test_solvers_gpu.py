"""GPU parity of the solver path (through the C ABI / drop-in API) against the reference's outputs
(tests/golden, made by the unmodified reference) and the oracle run live on small sizes.

Tolerances (BASELINE.json north_star): eigenvalues 1e-8 relative (absolute 1e-8 for lambda_0 ~ 0);
eigenvectors: sine of the largest principal angle per eigenvalue cluster <= 1e-6 (clusters =
relative gap < 1e-6, the cluster cut by k excluded); linear solves: 1e-8 relative to the solution
scale (the reference's own batch-vs-single tests use rtol 1e-6 / atol 1e-9).
"""

import numpy as np
import pytest
from conftest import golden_csc, golden_mesh, load_golden

pytestmark = pytest.mark.gpu

from oracle import diffgeo as odg  # noqa: E402
from oracle import fem as ofem  # noqa: E402
from oracle import solve as osolve  # noqa: E402

EV_RTOL = 1e-8


def check_evals(ev, ref, atol0=1e-8):
    ev, ref = np.asarray(ev), np.asarray(ref)
    assert ev.shape == ref.shape
    assert np.all(np.diff(ev) >= -1e-12), "ascending"
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    small = np.abs(ref) < 1e-3 * np.abs(ref).max()
    assert np.all(np.abs(ev - ref)[~small] <= EV_RTOL * np.abs(ref)[~small]), np.abs(ev - ref) / scale
    assert np.all(np.abs(ev - ref)[small] <= atol0)


def clusters(ev, rel=1e-6):
    out, cur = [], [0]
    for i in range(1, len(ev)):
        if abs(ev[i] - ev[i - 1]) <= rel * max(abs(ev[i]), 1e-3 * abs(ev[-1])):
            cur.append(i)
        else:
            out.append(cur)
            cur = [i]
    out.append(cur)
    return out


def max_sin_angle(x, xref, b, ev, drop_last=True):
    cl = clusters(ev)
    if drop_last:
        cl = cl[:-1]  # the cluster cut by k
    worst = 0.0
    for idx in cl:
        q, qr = x[:, idx], xref[:, idx]
        resid = qr - q @ (q.T @ (b @ qr))
        nrm = np.sqrt(np.abs(np.diag(resid.T @ (b @ resid))))
        worst = max(worst, nrm.max())
    return worst


@pytest.mark.parametrize("m", [1, 2, 3, 8, 13, 32, 50, 64, 96])
def test_spmm_matches_scipy(golden, m):
    from lapy_b200 import _lib

    g = golden("ico5")
    a = golden_csc(g, "A")
    ctx = _lib.default_context()
    da = _lib.DeviceMatrix.from_scipy(ctx, a)
    x = np.random.default_rng(m).standard_normal((a.shape[0], m))
    y = _lib.spmm(ctx, da, x)
    ref = a @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    y1 = _lib.spmm(ctx, da, x[:, 0])
    assert y1.shape == (a.shape[0],) and np.abs(y1 - ref[:, 0]).max() <= 1e-13 * np.abs(ref).max()
    bl = golden_csc(g, "B_lump")
    dl = _lib.DeviceMatrix.from_scipy(ctx, bl)
    assert np.abs(_lib.spmm(ctx, dl, x) - bl @ x).max() <= 1e-16


@pytest.mark.parametrize("maker,m", [("ico6", 64), ("ico6", 24), ("ico6", 128), ("cube15", 64), ("cube15", 8)])
def test_spmm_kernel_forms_agree(maker, m):
    """The strip-staged SpMM (default) against the row-wise kernel that test_spmm_matches_scipy pins, for all
    five epilogue modes: bit-identical in double (same summation order); the single-precision strip kernel of
    the preconditioner's multigrid cycle within float rounding of the double result."""
    from lapy_b200 import _lib
    from lapy_b200 import mesh as M

    mesh = M.icosphere(6) if maker == "ico6" else M.cube_tets(15)
    ctx = _lib.default_context()
    dm = _lib.DeviceMesh(ctx, mesh.v, mesh.t)
    a, _b = _lib.assemble(ctx, dm, _lib.FEM_TETRA if mesh.t.shape[1] == 4 else _lib.FEM_TRIA, False)
    errs = _lib.spmm_selftest(ctx, a, m)
    assert np.all(errs[:5] == 0.0), errs
    assert np.all(errs[5:] < 5e-6), errs


@pytest.mark.parametrize("name,k,lump", [("cubeTria", 10, False), ("squareMesh", 10, True), ("ico3", 20, False),
                                           ("cube9", 12, False), ("cubeTetra", 10, False), ("torus", 10, False)])  # fmt: skip
def test_eigs_vs_reference_golden(golden, name, k, lump):
    import lapy_b200

    g = golden(name)
    mesh = golden_mesh(g)
    fem = lapy_b200.Solver(mesh, lump=lump)
    ev, evec = fem.eigs(k=k)
    check_evals(ev, g["evals"], atol0=1e-8)
    assert evec.shape == (mesh.v.shape[0], k) and evec.flags.c_contiguous
    b = fem.mass
    np.testing.assert_allclose(evec.T @ (b @ evec), np.eye(k), atol=1e-9)
    r = fem.stiffness @ evec - (b @ evec) * ev
    assert np.abs(r).max() <= 1e-7 * max(1.0, np.abs(ev).max()) * np.abs(b @ evec).max()
    if "evecs" in g:
        assert max_sin_angle(evec, g["evecs"], b, g["evals"]) <= 1e-6


def test_reference_ev_files(golden):
    """data/cubeTria.ev, data/cubeTetra.ev as asserted by the reference's
    test_visualization_meshes.py:77,119 (k=3, rel 1e-5 / abs 1e-4) - here at 1e-8."""
    import lapy_b200

    for name in ("cubeTria", "cubeTetra"):
        g = golden(name)
        ev, evec = lapy_b200.Solver(golden_mesh(g)).eigs(k=3)
        check_evals(ev, g["ev_file"], atol0=1e-8)


def test_shapedna_ico5_k50(golden):
    import lapy_b200
    from lapy_b200.shapedna import compute_shapedna

    g = golden("ico5")
    mesh = golden_mesh(g)
    sd = compute_shapedna(mesh, k=50)
    check_evals(sd["Eigenvalues"], g["shapedna_evals"])
    meta = [sd[k] for k in ("Refine", "Degree", "Dimension", "Elements", "DoF", "NumEW")]
    np.testing.assert_array_equal(meta, g["shapedna_meta"])
    # eigenvectors against the oracle run live (2 s): per-cluster subspace angle
    a, b = ofem.fem(mesh)
    ev_ref, x_ref = osolve.eigs(a, b, k=50)
    assert max_sin_angle(sd["Eigenvectors"], x_ref, b, ev_ref) <= 1e-6
    lump = lapy_b200.Solver(mesh, lump=True).eigs(k=50)[0]
    assert abs(lump[1] - sd["Eigenvalues"][1]) < 1e-2


@pytest.mark.parametrize("key,maker", [("ico6_k50", lambda M: M.icosphere(6)), ("ico7_k50", lambda M: M.icosphere(7)),
                                       ("ico8_k50", lambda M: M.icosphere(8)), ("ico9_k50", lambda M: M.icosphere(9)),
                                       ("cube21_k50", lambda M: M.cube_tets(21)), ("cube31_k50", lambda M: M.cube_tets(31))])  # fmt: skip
def test_k50_spectra_vs_reference(key, maker):
    """k=50 spectra of the unmodified reference (tests/golden/spectra.npz; level 8 / 9 = BASELINE.md
    §5.2, level 9 is the north-star configuration: 2,621,442 vertices, 1e-8 relative)."""
    import lapy_b200
    from lapy_b200 import mesh as M

    ref = load_golden("spectra")[key]
    mesh = maker(M)
    fem = lapy_b200.Solver(mesh)
    ev, evec = fem.eigs(k=50)
    check_evals(ev, ref)
    # size-independent properties of the eigenvectors, on the device-assembled matrices
    b = fem.mass
    assert evec.shape == (mesh.v.shape[0], 50)
    np.testing.assert_allclose(evec.T @ (b @ evec), np.eye(50), atol=1e-9)
    r = fem.stiffness @ evec - (b @ evec) * ev
    assert np.abs(r).max() <= 1e-7 * np.abs(ev).max() * np.abs(b @ evec).max()


def test_k50_lumped_and_small_k():
    import lapy_b200
    from lapy_b200 import mesh as M

    ref = load_golden("spectra")["ico6_k50_lump"]
    fem = lapy_b200.Solver(M.icosphere(6), lump=True)
    check_evals(fem.eigs(k=50)[0], ref)
    check_evals(fem.eigs(k=3)[0], ref[:3])
    check_evals(fem.eigensystem(k=11)[0], ref[:11])
    # eigenvalues only (what batched_shapedna asks for): same spectrum, no eigenvector array
    ev, vec = fem.eigs(k=50, vectors=False)
    assert vec is None
    check_evals(ev, ref)
    # the persistent work blocks can be handed back to the pool; the next solve allocates again
    fem._ctx.release_workspace()
    check_evals(fem.eigs(k=50)[0], ref)
    tiny = lapy_b200.Solver(M.icosphere(1), lump=True)  # dense path
    ev_t, vec_t = tiny.eigs(k=5, vectors=False)
    assert vec_t is None and np.allclose(ev_t, tiny.eigs(k=5)[0], rtol=0, atol=1e-12)


def test_eigs_tiny_mesh_and_errors():
    import lapy_b200
    from lapy_b200 import mesh as M

    ico = M.icosphere(0)
    a, b = ofem.fem(ico)
    ev_ref, _ = osolve.eigs(a, b, k=3)
    fem = lapy_b200.Solver(ico)
    ev, evec = fem.eigs(k=3)
    check_evals(ev, ev_ref, atol0=1e-6)
    with pytest.raises(TypeError):
        fem.eigs(k=12)
    with pytest.raises(NotImplementedError):
        fem.eigs(k=3, sigma=5.0)


@pytest.mark.parametrize("name", ["cubeTria", "squareMesh", "cubeTetra", "ico3", "cube9", "ico5"])
def test_heat_geodesic_grad_div_vs_golden(golden, name):
    from lapy_b200 import diffgeo, heat

    g = golden(name)
    mesh = golden_mesh(g)
    seeds = g["heat_seeds"]
    u = heat.diffusion(mesh, seeds, m=1.0)
    ref_u = g["heat_u"]
    assert u.shape == ref_u.shape
    assert np.abs(u - ref_u).max() <= 1e-8 * np.abs(ref_u).max()
    # the heat method needs COMPONENTWISE relative accuracy (u spans 30-150 orders of magnitude and
    # compute_geodesic_f normalises its gradient): the sparse LU of the reference has it, so must we
    nz = np.abs(ref_u) > 1e-280
    assert np.all(np.abs(u - ref_u)[nz] <= 1e-7 * np.abs(ref_u)[nz]), (np.abs(u - ref_u)[nz] / np.abs(ref_u)[nz]).max()
    geo_own = diffgeo.compute_geodesic_f(mesh, u)  # geodesics from OUR heat solution
    assert np.abs(geo_own - g["geodesic"]).max() <= (1e-6 if mesh.v.dtype == np.float64 else 5e-6) * g["geodesic"].max()
    # gradient: same operation order as the reference -> bit-identical
    np.testing.assert_array_equal(diffgeo.compute_gradient(mesh, g["f"][:, 0]), g["grad_1d"])
    np.testing.assert_array_equal(diffgeo.compute_gradient(mesh, g["f"]), g["grad_2d"])
    # divergence: per-vertex summation order differs from SciPy's -> 1e-12 of the addends' scale
    for key_in, key_out in (("grad_1d", "div_1d"), ("grad_2d", "div_2d")):
        d = diffgeo.compute_divergence(mesh, g[key_in])
        assert d.shape == g[key_out].shape
        assert np.abs(d - g[key_out]).max() <= 1e-12 * np.abs(g[key_in]).max() * np.abs(mesh.v).max() * 8
    # float32 meshes: the "singular" Poisson operator has lambda_0 ~ -4e-6, the reference's LU
    # solution carries that noise -> 5e-6; float64 meshes agree to 1e-8
    gtol = 1e-8 if mesh.v.dtype == np.float64 else 5e-6
    geo = diffgeo.compute_geodesic_f(mesh, g["heat_u"])
    assert geo.shape == g["geodesic"].shape
    assert np.abs(geo - g["geodesic"]).max() <= gtol * g["geodesic"].max()
    geo2 = diffgeo.compute_geodesic_f(mesh, np.column_stack((g["heat_u"], g["f"][:, 0])))
    assert np.abs(geo2 - g["geodesic_2d"]).max() <= gtol * g["geodesic_2d"].max()


def test_heat_geodesic_poisson_renumbered_path_vs_oracle():
    """Meshes >= 20000 vertices are solved in the Morton-cell numbering (lb_solve); the answers must
    still come back in the caller's order and match the SuperLU oracle: heat componentwise, geodesics,
    Poisson with Dirichlet data and pure Neumann."""
    import lapy_b200
    from lapy_b200 import diffgeo, heat, mesh as M
    from oracle import fem as ofem, solve as osolve

    mesh = M.icosphere(6)  # 40962 vertices, hierarchical (poor-locality) vertex order
    seeds = [0, 17, 40000]
    u = heat.diffusion(mesh, seeds, m=1.0)
    ref_u = osolve.diffusion(mesh, seeds, m=1.0)
    nz = np.abs(ref_u) > 1e-280
    assert np.all(np.abs(u - ref_u)[nz] <= 1e-7 * np.abs(ref_u)[nz]), (np.abs(u - ref_u)[nz] / np.abs(ref_u)[nz]).max()
    um = heat.diffusion(mesh, [[0], [5, 6, 7]], m=4.0)  # two seed sets -> two columns
    ref_um = osolve.diffusion(mesh, [[0], [5, 6, 7]], m=4.0)
    nz = np.abs(ref_um) > 1e-280
    assert um.shape == ref_um.shape
    assert np.all(np.abs(um - ref_um)[nz] <= 1e-7 * np.abs(ref_um)[nz])
    geo = diffgeo.compute_geodesic_f(mesh, u)
    ref_geo = osolve.geodesic_f(mesh, ref_u)
    assert np.abs(geo - ref_geo).max() <= 1e-6 * ref_geo.max()
    fem = lapy_b200.Solver(mesh, lump=True)
    a, b = ofem.fem(mesh, lump=True)
    rng = np.random.default_rng(3)
    h = rng.standard_normal((len(mesh.v), 2))
    didx = np.array([3, 1000, 20000, 40961])
    dval = np.array([0.5, -1.0, 2.0, 0.0])
    x = fem.poisson(h, dtup=(didx, dval))
    ref = osolve.poisson(a, b, h, dtup=(didx, dval))
    assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max()
    np.testing.assert_array_equal(x[didx], np.column_stack((dval, dval)))
    h0 = h[:, 0] - (b @ h[:, 0]).sum() / b.sum()  # compatible right-hand side for the singular system
    x0 = fem.poisson(h0)
    ref0 = osolve.poisson(a, b, h0)
    assert np.abs((x0 - x0.mean()) - (ref0 - ref0.mean())).max() <= 1e-7 * np.abs(ref0 - ref0.mean()).max()


def test_reference_geodesic_expected_outcomes(golden):
    """test_TriaMesh_Geodesics.py:185 / expected_outcomes.json: max geodesic 0.60497826 (rtol 1e-5)."""
    from lapy_b200 import Solver, diffgeo, heat

    g = golden("squareMesh")
    mesh = golden_mesh(g)
    u = heat.diffusion(mesh, g["boundary"], m=1)
    gf = diffgeo.compute_geodesic_f(mesh, u)
    assert np.isclose(gf.max(), 0.60497826, rtol=1e-5)
    fem = Solver(mesh, lump=True)
    assert np.isclose(fem.mass.sum(), 1.0, rtol=1e-6)
    ad = fem.stiffness.toarray()
    assert (ad == ad.T).all()
    assert fem.stiffness.getformat() == "csc"
    h = -fem.stiffness  # callers negate / combine the matrices (test_TriaMesh_Geodesics.py:152-177)
    assert h.getformat() == "csc"


@pytest.mark.parametrize("name", ["cubeTria", "squareMesh", "ico3", "cube9"])
def test_poisson_vs_golden(golden, name):
    import lapy_b200

    g = golden(name)
    fem = lapy_b200.Solver(golden_mesh(g), lump=bool(g["poisson_lump"]))
    h = g["poisson_h"]
    dt = (g["poisson_didx"], g["poisson_dval"])
    nt = (g["poisson_nidx"], g["poisson_nval"])

    def close(x, ref, rtol=1e-8):
        assert x.shape == ref.shape
        assert np.abs(x - ref).max() <= rtol * max(np.abs(ref).max(), 1e-12), np.abs(x - ref).max()

    # float32 meshes: A 1 != 0 by ~1e-7 (lambda_0 ~ -4e-6), so the reference's LU solution of the
    # "singular" system carries an amplified near-constant component: compare at 1e-6 there
    ntol = 1e-8 if g["v"].dtype == np.float64 else 2e-6

    close(fem.poisson(h[:, 0], dtup=dt), g["poisson_dirichlet_1d"])
    close(fem.poisson(h, dtup=dt), g["poisson_dirichlet_2d"])
    close(fem.poisson(0.0, dtup=dt), g["poisson_laplace_dirichlet"])
    close(fem.poisson(h[:, 0], dtup=dt, ntup=nt), g["poisson_neumann_dirichlet"])
    x = fem.poisson(h)  # pure Neumann: defined up to a constant per column
    r = g["poisson_2d"]
    close(x - x.mean(0), r - r.mean(0), ntol)
    x1 = fem.poisson(h[:, 0])
    assert x1.ndim == 1
    close(x1 - x1.mean(), r[:, 0] - r[:, 0].mean(), ntol)


# ---- the reference's own solver / heat tests, re-pointed at the drop-in -------------------------
@pytest.fixture
def square(golden):
    return golden_mesh(golden("squareMesh"))


def test_ref_poisson_scalar_and_1d_return_1d(square):  # test_solver.py:15-22
    from lapy_b200 import Solver

    fem = Solver(square, lump=True)
    _, evec = fem.eigs(k=3)
    assert fem.poisson(0.0).ndim == 1
    assert fem.poisson(evec[:, 1]).ndim == 1


def test_ref_poisson_2d_rhs_matches_1d(square):  # test_solver.py:25-40, :43-59
    from lapy_b200 import Solver

    fem = Solver(square, lump=True)
    _, evec = fem.eigs(k=5)
    rhs = evec[:, 1:5]
    xb = fem.poisson(rhs)
    assert xb.shape == (len(square.v), 4)
    for k in range(4):
        np.testing.assert_allclose(xb[:, k], fem.poisson(rhs[:, k]), rtol=1e-6, atol=1e-9)
    dtup = (np.array([0, 1]), np.array([0.0, 0.0]))
    xb = fem.poisson(rhs[:, :3], dtup=dtup)
    for k in range(3):
        np.testing.assert_allclose(xb[:, k], fem.poisson(rhs[:, k], dtup=dtup), rtol=1e-6, atol=1e-9)


def test_ref_diffusion_shapes_multi_and_errors(golden, square):  # test_heat.py:20-72
    from lapy_b200.heat import diffusion

    g = golden("squareMesh")
    bvert = g["boundary"]
    n = len(square.v)
    assert diffusion(square, bvert, m=1).shape == (n,)
    assert diffusion(square, bvert.tolist(), m=1).shape == (n,)
    assert diffusion(square, 0, m=1).shape == (n,)
    seeds = [bvert, np.array([0]), np.array([1, 2])]
    ub = diffusion(square, seeds, m=1)
    assert ub.shape == (n, 3)
    np.testing.assert_allclose(ub, g["heat_multi"], rtol=1e-6, atol=1e-9)
    for k, s in enumerate(seeds):
        np.testing.assert_allclose(ub[:, k], diffusion(square, s, m=1), rtol=1e-6, atol=1e-9)
    assert diffusion(square, [[0, 1], [2]], m=1).shape == (n, 2)
    with pytest.raises(ValueError, match="out-of-range"):
        diffusion(square, np.array([n]), m=1)
    with pytest.raises(ValueError, match="out-of-range"):
        diffusion(square, [np.array([0]), np.array([n])], m=1)


def test_ref_laplace_identity(golden):  # test_diffgeo.py:95-110: -div(grad f_k) = lambda_k B f_k
    import lapy_b200
    from lapy_b200 import diffgeo

    mesh = golden_mesh(golden("ico5"))
    fem = lapy_b200.Solver(mesh, lump=True)
    ev, evec = fem.eigs(k=4)
    for k in range(1, 4):
        lhs = -diffgeo.compute_divergence(mesh, diffgeo.compute_gradient(mesh, evec[:, k]))
        rhs = ev[k] * (fem.mass @ evec[:, k])
        np.testing.assert_allclose(lhs, rhs, rtol=1e-3, atol=1e-6 * np.abs(rhs).max())


def test_user_assigned_mass_and_errors(golden):
    import lapy_b200
    from scipy import sparse

    g = golden("ico3")
    fem = lapy_b200.Solver(golden_mesh(g), lump=True)
    n = fem.stiffness.shape[0]
    fem.mass = sparse.eye(n, dtype=np.float64)  # lapy/diffgeo.py:149
    a = golden_csc(g, "A")
    h = g["poisson_h"][:, 0]
    h = h - h.mean()
    x = fem.poisson(h)
    ref = osolve.poisson(a, sparse.eye(n, format="csc"), h)
    assert np.abs((x - x.mean()) - (ref - ref.mean())).max() <= 1e-8 * np.abs(ref).max()
    with pytest.raises(ValueError):
        fem.poisson(np.zeros(n + 1))
    with pytest.raises(ValueError, match="unique"):
        fem.poisson(h, dtup=(np.array([0, 0]), np.array([1.0, 1.0])))
    with pytest.raises(ValueError):
        fem.poisson(h, dtup=(np.array([0, 1]),))


@pytest.mark.parametrize("n,p,q", [(1000, 8, 8), (4097, 64, 64), (10001, 192, 192), (5003, 83, 21), (20000, 128, 60), (777, 1, 5)])
def test_dmma_block_products_match_numpy(n, p, q):
    """Hand-written DMMA Gram / update kernels vs NumPy (fp64, different summation order)."""
    from lapy_b200 import _lib

    ctx = _lib.default_context()
    rng = np.random.default_rng(n + p + q)
    x, y = rng.standard_normal((n, p)), rng.standard_normal((n, q))
    c = _lib.block_gram(ctx, x, y)
    ref = x.T @ y
    assert np.abs(c - ref).max() <= 1e-12 * np.sqrt(n) * 10
    cm = rng.standard_normal((p, q))
    y0 = rng.standard_normal((n, q))
    out = _lib.block_update(ctx, x, cm, alpha=0.75, beta=-1.5, y=y0)
    ref = 0.75 * x @ cm - 1.5 * y0
    assert np.abs(out - ref).max() <= 1e-12 * p
    out = _lib.block_update(ctx, x, cm)
    assert np.abs(out - x @ cm).max() <= 1e-12 * p


def test_level7_heat_geodesic_reference_values():
    """values the unmodified reference produced on the level-7 icosphere
    (163,842 vertices): t, u[0], sum(u) of heat.diffusion(T,[0],m=1) and max / mean of
    compute_geodesic_f(T,u).  The mesh is float64 with >= 100k vertices, so ``t`` comes from the
    device edge-length kernel (lb_avg_edge_length), checked against the host formula as well."""
    from lapy_b200 import _lib, diffgeo, heat
    from lapy_b200 import mesh as M
    from lapy_b200.solver import Solver

    mesh = M.icosphere(7)
    fem = Solver(mesh, lump=True)
    h_dev = _lib.avg_edge_length(fem._ctx, fem._mesh, fem._device("a"))
    assert abs(h_dev - mesh.avg_edge_length()) <= 1e-13 * h_dev
    assert abs(h_dev**2 - 8.916850393006576e-05) <= 1e-12 * 8.916850393006576e-05
    u = heat.diffusion(mesh, [0], m=1.0)
    assert abs(u[0] - 3325.8603222719325) <= 1e-8 * 3325.8603222719325
    assert abs(u.sum() - 14606.340887242888) <= 1e-8 * 14606.340887242888
    g = diffgeo.compute_geodesic_f(mesh, u)
    assert abs(g.max() - 3.1302828669840617) <= 1e-6 * np.pi
    assert abs(g.mean() - 1.5653885315086715) <= 1e-6 * np.pi


def test_assigned_nonsingular_stiffness_is_not_projected(golden):
    """``fem.stiffness = A + c*B`` (screened Poisson) is nonsingular; without Dirichlet data
    the solver must return K^-1 b like the reference's splu, not a zero-mean projection."""
    import lapy_b200
    from scipy.sparse.linalg import splu

    g = golden("ico5")
    fem = lapy_b200.Solver(golden_mesh(g), lump=True)
    a, b = golden_csc(g, "A"), golden_csc(g, "B_lump")
    k = (a + 0.5 * b).tocsc()
    fem.stiffness = k
    h = np.random.default_rng(1).standard_normal(a.shape[0]) + 3.0  # nonzero mean
    x = fem.poisson(h)
    ref = splu(k).solve(b @ h)
    assert abs(ref.mean()) > 1.0  # the projected answer would have mean 0
    assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max()


def test_in_place_mesh_edit_is_seen(golden):
    """the reference reads the current host arrays on every call; an in-place edit of
    ``mesh.v`` between two Solver constructions must not be served from a stale device copy."""
    import lapy_b200

    g = golden("ico3")
    mesh = golden_mesh(g)
    m0 = lapy_b200.Solver(mesh, lump=True).mass.sum()
    mesh.v *= 2.0
    m1 = lapy_b200.Solver(mesh, lump=True).mass.sum()
    assert abs(m1 - 4.0 * m0) <= 1e-12 * m1
    mesh.t[:, [1, 2]] = mesh.t[:, [2, 1]]  # flips the orientation: same matrices, different triplet order
    a2 = lapy_b200.Solver(mesh, lump=True).stiffness
    assert abs(a2 - 1.0 * golden_csc(g, "A")).max() < 1e-12
    # read-only arrays opt in to the device-side cache
    mesh.v.flags.writeable = False
    mesh.t.flags.writeable = False
    s1 = lapy_b200.Solver(mesh)
    s2 = lapy_b200.Solver(mesh)
    assert s1._mesh is s2._mesh
