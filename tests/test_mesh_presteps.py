"""CPU: the mesh-side pre-steps of lapy_b200.mesh against outputs of the unmodified
reference (tests/golden/frows.npz, tools/make_golden_frows.py; ico3 / torus goldens of tools/make_golden.py).
These are host-side NumPy restatements that keep the reference's operation order: results are expected
bit-identical, asserted to 1e-13."""

import numpy as np
import pytest
from conftest import golden_mesh, load_golden

from lapy_b200 import mesh as M


@pytest.fixture(scope="module")
def fr():
    return load_golden("frows")


def test_triamesh_measures_edges_normals(fr):
    m = M.TriaMesh(fr["ell_v"], fr["ell_t"])
    assert m.is_closed() and m.is_oriented() and m.is_manifold()
    assert m.avg_edge_length() == pytest.approx(float(fr["ell_avg_edge"]), rel=1e-14)
    assert m.area() == pytest.approx(float(fr["ell_area"]), rel=1e-14)
    assert m.volume() == pytest.approx(float(fr["ell_volume"]), rel=1e-14)
    vids, tids = m.edges()
    np.testing.assert_array_equal(vids, fr["ell_edges_vids"])
    np.testing.assert_array_equal(tids, fr["ell_edges_tids"])
    np.testing.assert_allclose(m.vertex_normals(), fr["ell_vertex_normals"], rtol=0, atol=1e-13)
    c, area = m.centroid()
    m.normalize_()
    c2, area2 = m.centroid()
    assert abs(area2 - 1.0) < 1e-12 and np.abs(c2).max() < 1e-12


def test_curvature_and_curvature_tria(fr):
    m = M.TriaMesh(fr["ell_v"], fr["ell_t"])
    for k, arr in zip(("umin", "umax", "cmin", "cmax", "cmean", "cgauss", "normals"), m.curvature(smoothit=3)):
        np.testing.assert_allclose(arr, fr["ell_curv_" + k], rtol=0, atol=1e-12, err_msg=k)
    for k, arr in zip(("u1", "u2", "c1", "c2"), m.curvature_tria(smoothit=10)):
        np.testing.assert_allclose(arr, fr["ell_curvtria_" + k], rtol=0, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("name,aniso,smooth", [("ico3", (2.0, 5.0), 3), ("torus", (1.0, 10.0), 2)])
def test_aniso_inputs_from_own_curvature(name, aniso, smooth):
    """The (u1, u2, aniso_mat) triple Solver.__init__ builds (lapy/solver.py:73-91) from THIS package's
    curvature_tria equals the one the reference built from its own."""
    g = load_golden(name)
    m = golden_mesh(g)
    u1, u2, c1, c2 = m.curvature_tria(smoothit=smooth)
    am = np.empty((m.t.shape[0], 2))
    am[:, 1] = np.exp(-aniso[1] * np.abs(c1))
    am[:, 0] = np.exp(-aniso[0] * np.abs(c2))
    np.testing.assert_allclose(u1, g["aniso_u1"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(u2, g["aniso_u2"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(am, g["aniso_mat"], rtol=0, atol=1e-12)


def test_tet_adjacency_and_edge_length(fr):
    from scipy import sparse

    t = M.cube_tets(6)
    a = t.adj_sym
    ref = sparse.csc_matrix((fr["tet_adj_data"], fr["tet_adj_indices"], fr["tet_adj_indptr"]), shape=a.shape)
    assert (a != ref).nnz == 0
    assert t.avg_edge_length() == pytest.approx(float(fr["tet_avg_edge"]), rel=1e-14)


def test_open_mesh_edges_and_errors():
    g = load_golden("squareMesh")
    m = golden_mesh(g)
    assert not m.is_closed()
    vids, tids = m.edges()  # inner edges only
    assert len(vids) == (m.adj_sym.data == 2).sum() // 2
    with pytest.raises(ValueError, match="closed"):
        m.volume()
    flipped = M.TriaMesh(m.v, np.vstack([m.t[:1, ::-1], m.t[1:]]))
    assert not flipped.is_oriented()
    with pytest.raises(ValueError, match="oriented"):
        flipped.edges()
