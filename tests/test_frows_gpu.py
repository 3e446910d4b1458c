"""GPU parity of the callers either side of the hot path: anisotropic operator through
the public constructor on THIS package's TriaMesh (curvature_tria), tria_compute_divergence2,
tria_compute_rotated_f, tria_mean_curvature_flow, tria_spherical_project - against outputs of the
unmodified reference (tests/golden/frows.npz, tools/make_golden_frows.py)."""

import numpy as np
import pytest
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fr():
    return load_golden("frows")


@pytest.fixture
def ell(fr):
    from lapy_b200.mesh import TriaMesh

    return TriaMesh(fr["ell_v"], fr["ell_t"])


def test_aniso_solver_on_own_mesh(fr, ell):
    import lapy_b200
    from lapy_b200 import heat

    fem = lapy_b200.Solver(ell, aniso=(1.0, 4.0), aniso_smooth=5)
    a = fem.stiffness
    np.testing.assert_array_equal(a.indptr, fr["ell_aniso_Aindptr"])
    np.testing.assert_array_equal(a.indices, fr["ell_aniso_Aindices"])
    scale = np.abs(fr["ell_aniso_Adata"]).max()
    assert np.abs(a.data - fr["ell_aniso_Adata"]).max() <= 1e-12 * scale
    ev = fem.eigs(k=8)[0]
    ref = fr["ell_aniso_evals"]
    assert np.all(np.abs(ev[1:] - ref[1:]) <= 1e-8 * np.abs(ref[1:])) and abs(ev[0]) < 1e-8
    with pytest.raises(ValueError, match="length 2"):
        lapy_b200.Solver(ell, aniso=(1.0, 2.0, 3.0))
    u = heat.diffusion(ell, [0, 50], m=1.0, aniso=2.0)
    ref_u = fr["ell_heat_aniso"]
    nz = np.abs(ref_u) > 1e-280
    assert np.all(np.abs(u - ref_u)[nz] <= 1e-7 * np.abs(ref_u)[nz])


def test_divergence2_and_rotated_f(fr, ell):
    from lapy_b200 import diffgeo

    f = fr["ell_f"]
    g1, g2 = diffgeo.tria_compute_gradient(ell, f[:, 0]), diffgeo.tria_compute_gradient(ell, f)
    for got, ref in ((diffgeo.tria_compute_divergence2(ell, g1), fr["ell_div2_1d"]),
                     (diffgeo.tria_compute_divergence2(ell, g2), fr["ell_div2_2d"])):  # fmt: skip
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(g2).max() * np.abs(ell.v).max() * 8
    # the flux form equals the cotangent form up to rounding
    assert np.abs(diffgeo.tria_compute_divergence2(ell, g1) - diffgeo.tria_compute_divergence(ell, g1)).max() < 1e-10
    r1 = diffgeo.tria_compute_rotated_f(ell, f[:, 0])
    r2 = diffgeo.tria_compute_rotated_f(ell, f)
    assert r1.shape == fr["ell_rot_1d"].shape and r2.shape == fr["ell_rot_2d"].shape
    # the rotated gradient of a function is divergence free: both sides are rounding noise (1e-16) around 0
    # for this input - compared on the scale of f - plus a field that is NOT a gradient below
    scale = np.abs(f).max()
    assert np.abs(r1 - fr["ell_rot_1d"]).max() <= 1e-10 * scale
    assert np.abs(r2 - fr["ell_rot_2d"]).max() <= 1e-10 * scale
    assert r1[0] == 0.0  # the pinned vertex
    # open mesh (float32 vertices): boundary flux makes the result non-trivial (f = x -> ~ y)
    from conftest import golden_mesh, load_golden

    sq = golden_mesh(load_golden("squareMesh"))
    fs = fr["sq_f"]
    for got, ref in ((diffgeo.tria_compute_rotated_f(sq, fs[:, 0]), fr["sq_rot_1d"]),
                     (diffgeo.tria_compute_rotated_f(sq, fs), fr["sq_rot_2d"])):  # fmt: skip
        assert got.shape == ref.shape and np.abs(ref).max() > 0.1
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()  # float32 mesh, see test_poisson_vs_golden
    d2 = diffgeo.tria_compute_divergence2(sq, diffgeo.tria_compute_gradient(sq, fs[:, 1]))
    assert np.abs(d2 - fr["sq_div2"]).max() <= 1e-12 * 8


def test_mean_curvature_flow(fr, ell):
    from lapy_b200 import diffgeo

    out = diffgeo.tria_mean_curvature_flow(ell, max_iter=3)
    assert type(out).__name__ == "TriaMesh" and out is not ell
    assert np.abs(out.v - fr["ell_mcf3_v"]).max() <= 1e-8 * np.abs(fr["ell_mcf3_v"]).max()
    assert abs(out.area() - 1.0) < 1e-10
    np.testing.assert_array_equal(ell.v, fr["ell_v"])  # the input mesh is not modified
    out = diffgeo.tria_mean_curvature_flow(ell, max_iter=8, step=0.5)
    assert np.abs(out.v - fr["ell_mcf8_step05_v"]).max() <= 1e-8 * np.abs(fr["ell_mcf8_step05_v"]).max()


def test_spherical_project(fr, ell, golden):
    from conftest import golden_mesh

    from lapy_b200 import diffgeo

    out = diffgeo.tria_spherical_project(ell, flow_iter=3)
    ref = fr["ell_sphere_v"]
    np.testing.assert_allclose(np.sqrt((out.v**2).sum(1)), 100.0, rtol=1e-12)
    assert np.abs(out.v - ref).max() <= 1e-5 * 100.0
    with pytest.raises(ValueError, match="closed"):
        diffgeo.tria_spherical_project(golden_mesh(golden("squareMesh")))
