"""GPU parity: CUDA assembly (through the C ABI / Solver drop-in) vs the oracle and the goldens.

Bars (BASELINE.json north_star): indptr / indices bit-exact to SciPy's COO->CSC; values
within 1e-12 relative.  We check more: values are BIT-IDENTICAL to the oracle's restatement of
the kernel's summation order (oracle.sparse_build.coo_to_csc_sequential), and within
1e-12 * (sum of |addends|) of SciPy's own result (equal entries where SciPy's order coincides).
"""

import numpy as np
import pytest
from conftest import golden_csc, golden_mesh

pytestmark = pytest.mark.gpu

from oracle import fem as ofem  # noqa: E402
from oracle.sparse_build import coo_to_csc_sequential  # noqa: E402


def _triplets(v, t, kind, aniso=None):
    """COO triplets (A, B_full, B_lump per corner) of the reference, from the oracle's locals."""
    if kind == "tet":
        (a12, a13, a14, a23, a24, a34), vol = ofem.tet_local(v, t)
        a11, a22 = -a12 - a13 - a14, -a12 - a23 - a24
        a33, a44 = -a13 - a23 - a34, -a14 - a24 - a34
        cols = (a12, a12, a23, a23, a13, a13, a14, a14, a24, a24, a34, a34, a11, a22, a33, a44)
        da, i, j = ofem._coo(t, ofem.TET_SLOTS, cols)
        da = da / 6.0
        bii, bij, bl = vol / 60.0, vol / 120.0, vol / 24.0
        db, _, _ = ofem._coo(t, ofem.TET_SLOTS, (bij,) * 12 + (bii,) * 4)
    else:
        if kind == "mass":
            p1, p2, p3 = (ofem._corner(v, t, c) for c in range(3))
            cr = ofem._cross(ofem._sub(p3, p2), ofem._sub(p1, p3))
            vol = ofem._clamp(0.5 * np.sqrt(ofem._dot(cr, cr)))
            a12 = a23 = a31 = np.zeros_like(vol)
            bii, bij, bl = vol / 6, vol / 12, vol / 3
        else:
            a12, a23, a31, vol = ofem.tria_local(v, t, aniso)
            bii, bij, bl = vol / 24, vol / 48, vol / 12
        cols = (a12, a12, a23, a23, a31, a31, -a12 - a31, -a12 - a23, -a31 - a23)
        da, i, j = ofem._coo(t, ofem.TRIA_SLOTS, cols)
        db, _, _ = ofem._coo(t, ofem.TRIA_SLOTS, (bij,) * 6 + (bii,) * 3)
    k = t.shape[1]
    il = np.column_stack([t[:, c] for c in range(k)]).reshape(-1)
    dl = np.column_stack([bl] * k).reshape(-1)
    f64 = np.float64
    return (da.astype(f64), i, j), (db.astype(f64), i, j), (dl.astype(f64), il, il)


def _check(m, trip, ref=None, name=""):
    dat, i, j = trip
    indptr, indices, data = coo_to_csc_sequential(dat, i, j)
    assert m.format == "csc" and m.indices.dtype == np.int32 and m.indptr.dtype == np.int32
    assert m.shape == (len(indptr) - 1,) * 2, name
    np.testing.assert_array_equal(m.indptr, indptr, err_msg=name)
    np.testing.assert_array_equal(m.indices, indices, err_msg=name)
    np.testing.assert_array_equal(m.data, data, err_msg=name + " (bitwise vs sequential model)")
    if ref is not None:
        ref = ref.tocsc()
        np.testing.assert_array_equal(m.indptr, ref.indptr)
        np.testing.assert_array_equal(m.indices, ref.indices)
        mag = coo_to_csc_sequential(np.abs(dat), i, j)[2]
        assert np.all(np.abs(m.data - ref.data) <= 1e-12 * mag), name + " (vs SciPy)"


def _kind(mesh):
    return "tet" if mesh.t.shape[1] == 4 else "tria"


GOLDEN = ["cubeTria", "squareMesh", "cubeTetra", "ico3", "torus", "ico5", "cube9", "degenerate"]


@pytest.mark.parametrize("name", GOLDEN)
def test_solver_matrices_vs_golden(golden, name):
    import lapy_b200

    g = golden(name)
    mesh = golden_mesh(g)
    ta, tb, tl = _triplets(mesh.v, mesh.t, _kind(mesh))
    fem = lapy_b200.Solver(mesh)
    _check(fem.stiffness, ta, golden_csc(g, "A"), name + " A")
    _check(fem.mass, tb, golden_csc(g, "B_full"), name + " B")
    fem = lapy_b200.Solver(mesh, lump=True)
    _check(fem.stiffness, ta, golden_csc(g, "A"), name + " A(lump)")
    _check(fem.mass, tl, golden_csc(g, "B_lump"), name + " Blump")


@pytest.mark.parametrize("name", ["cubeTria", "ico3", "degenerate"])
def test_fem_tria_mass_vs_golden(golden, name):
    import lapy_b200

    g = golden(name)
    mesh = golden_mesh(g)
    _, tb, tl = _triplets(mesh.v, mesh.t, "mass")
    _check(lapy_b200.Solver.fem_tria_mass(mesh), tb, golden_csc(g, "M_full"), name)
    _check(lapy_b200.Solver.fem_tria_mass(mesh, lump=True), tl, golden_csc(g, "M_lump"), name)


@pytest.mark.parametrize("name", ["ico3", "torus"])
def test_aniso_vs_golden(golden, name):
    import lapy_b200

    g = golden(name)
    mesh = golden_mesh(g)
    an = (g["aniso_u1"], g["aniso_u2"], g["aniso_mat"])
    ta, tb, tl = _triplets(mesh.v, mesh.t, "tria", an)
    a, b = lapy_b200.Solver._fem_tria_aniso(mesh, *an)
    _check(a, ta, golden_csc(g, "A_aniso"), name + " A")
    _check(b, tb, golden_csc(g, "B_aniso"), name + " B")
    a, b = lapy_b200.Solver._fem_tria_aniso(mesh, *an, lump=True)
    _check(b, tl, None, name + " Blump")


def test_static_entry_points_and_dtypes(golden):
    import lapy_b200
    from lapy_b200.mesh import TetMesh, TriaMesh

    g = golden("ico3")
    for tdt in (np.int32, np.int64, ">i4", np.uint32):
        mesh = TriaMesh(g["v"], g["t"].astype(tdt))
        a, b = lapy_b200.Solver._fem_tria(mesh)
        assert (a != golden_csc(g, "A")).nnz == 0 or abs(a - golden_csc(g, "A")).max() < 1e-14
        assert abs(b - golden_csc(g, "B_full")).max() < 1e-16
    # non-contiguous vertex view (TriaMesh may hold a transposed view)
    mesh = TriaMesh(g["v"], g["t"])
    mesh.v = np.asfortranarray(mesh.v)
    a, _ = lapy_b200.Solver._fem_tria(mesh)
    assert abs(a - golden_csc(g, "A")).max() < 1e-14
    g = golden("cube9")
    a, b = lapy_b200.Solver._fem_tetra(TetMesh(g["v"], g["t"]), lump=True)
    assert abs(b - golden_csc(g, "B_lump")).max() < 1e-18
    a32, _ = lapy_b200.Solver._fem_tetra(TetMesh(g["v"], g["t"]), dtype=np.float32)
    assert a32.dtype == np.float32


def test_errors():
    import lapy_b200
    from lapy_b200.mesh import TriaMesh

    class Quad:
        v = np.zeros((4, 3))
        t = np.zeros((1, 4), int)

    with pytest.raises(ValueError, match="unknown"):
        lapy_b200.Solver(Quad())
    m = TriaMesh(np.random.rand(5, 3), np.array([[0, 1, 2], [2, 3, 4]]))
    m.t = np.array([[0, 1, 7]])
    with pytest.raises(ValueError, match="Max index"):
        lapy_b200.Solver(m)


@pytest.mark.parametrize("level", [6, 7])
def test_icosphere_vs_oracle(level):
    import lapy_b200
    from lapy_b200 import mesh as M

    mesh = M.icosphere(level)
    a_ref, b_ref = ofem.fem(mesh)
    ta, tb, tl = _triplets(mesh.v, mesh.t, "tria")
    fem = lapy_b200.Solver(mesh)
    _check(fem.stiffness, ta, a_ref, "A")
    _check(fem.mass, tb, b_ref, "B")
    a = fem.stiffness
    # manifold triangle meshes: off-diagonals have 2 addends -> bitwise symmetric and equal to SciPy
    off = a.indices != np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    assert np.array_equal(a.data[off], a_ref.data[off])
    assert (a != a.T).nnz == 0


def test_tet_cube_vs_oracle():
    import lapy_b200
    from lapy_b200 import mesh as M

    mesh = M.cube_tets(25)
    a_ref, b_ref = ofem.fem(mesh)
    ta, tb, tl = _triplets(mesh.v, mesh.t, "tet")
    fem = lapy_b200.Solver(mesh)
    _check(fem.stiffness, ta, a_ref, "A")
    _check(fem.mass, tb, b_ref, "B")
    _, bl_ref = ofem.fem(mesh, lump=True)
    _check(lapy_b200.Solver(mesh, lump=True).mass, tl, bl_ref, "Blump")


def _tet_fan():
    """All 320 surface triangles of a level-2 icosphere joined to the centre: a vertex with 320
    incident tets and 162 neighbours (oversized row), surface vertices with 5-6."""
    from lapy_b200 import mesh as M

    ico = M.icosphere(2)
    v = np.vstack([ico.v, [[0.01, -0.02, 0.03]]])
    t = np.column_stack([ico.t, np.full(len(ico.t), len(ico.v))])
    return M.TetMesh(v, t)


def test_tet_fan_and_float32_rows_vs_oracle():
    """Tet rows: the fused kernels handle <= 32 incidences / <= 16 entries per row, everything else is
    flagged and done by the per-thread kernels: a vertex with 320 incident tets (162 neighbours), a
    float32 cube and the lumped mass must all reproduce the sequential model bit for bit."""
    import lapy_b200
    from lapy_b200 import mesh as M

    fan = _tet_fan()
    ta, tb, tl = _triplets(fan.v, fan.t, "tet")
    a_ref, b_ref = ofem.fem(fan)
    fem = lapy_b200.Solver(fan)
    _check(fem.stiffness, ta, a_ref, "fan A")
    _check(fem.mass, tb, b_ref, "fan B")
    _check(lapy_b200.Solver(fan, lump=True).mass, tl, None, "fan Blump")
    cube = M.cube_tets(13)
    cube32 = M.TetMesh(cube.v.astype(np.float32), cube.t)
    ta, tb, tl = _triplets(cube32.v, cube32.t, "tet")
    a_ref, b_ref = ofem.fem(cube32)
    fem = lapy_b200.Solver(cube32)
    _check(fem.stiffness, ta, a_ref, "cube32 A")
    _check(fem.mass, tb, b_ref, "cube32 B")


def test_triangle_fast_path_fallbacks_vs_oracle():
    """Triangle rows: the register fast path handles <= 8 incident triangles; rows with more, a
    non-manifold edge (three triangles on one edge: three addends, order matters) or a vertex repeated
    inside a triangle go through the flagged path.  Open mesh, float32, lumped mass as well."""
    import lapy_b200
    from lapy_b200 import mesh as M
    from lapy_b200.mesh import TriaMesh

    ico = M.icosphere(4)
    gx, gy = np.meshgrid(np.arange(40), np.arange(30), indexing="ij")
    gid = (gx * 30 + gy)[:-1, :-1].reshape(-1)
    square = TriaMesh(np.column_stack([gx.reshape(-1) * 0.1, gy.reshape(-1) * 0.13, np.zeros(gx.size)]),
                      np.vstack([np.column_stack([gid, gid + 30, gid + 31]), np.column_stack([gid, gid + 31, gid + 1])]))  # fmt: skip
    # valence-12 vertices: a 12-triangle fan glued into a strip
    n = 12
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    fan12 = TriaMesh(np.vstack([[0, 0, 0.3], np.column_stack([np.cos(ang), np.sin(ang), np.zeros(n)])]),
                     np.column_stack([np.zeros(n, int), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n]))  # fmt: skip
    # three triangles sharing the edge (0, 1) + one triangle with a repeated vertex
    book = TriaMesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, -1, 0.5], [0.3, 0.2, 1.0], [2, 2, 2]], float),
                    np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4], [2, 2, 5], [4, 1, 5]]))
    for name, mesh in (("ico4", ico), ("ico4_f32", TriaMesh(ico.v.astype(np.float32), ico.t)), ("square", square),
                       ("fan12", fan12), ("book", book)):  # fmt: skip
        ta, tb, tl = _triplets(mesh.v, mesh.t, "tria")
        a_ref, b_ref = ofem.fem(mesh)
        fem = lapy_b200.Solver(mesh)
        _check(fem.stiffness, ta, a_ref, name + " A")
        _check(fem.mass, tb, b_ref, name + " B")
        _, bl_ref = ofem.fem(mesh, lump=True)
        _check(lapy_b200.Solver(mesh, lump=True).mass, tl, bl_ref, name + " Blump")
        _, tbm, tlm = _triplets(mesh.v, mesh.t, "mass")
        _check(lapy_b200.Solver.fem_tria_mass(mesh), tbm, None, name + " M")
        _check(lapy_b200.Solver.fem_tria_mass(mesh, lump=True), tlm, None, name + " Mlump")


def test_high_valence_fan():
    """A vertex with 3000 incident triangles: exercises the oversized-block paths."""
    import lapy_b200
    from lapy_b200.mesh import TriaMesh

    n = 3000
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    v = np.vstack([[0, 0, 0.3], np.column_stack([np.cos(ang), np.sin(ang), np.zeros(n)])])
    t = np.column_stack([np.zeros(n, int), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n])
    mesh = TriaMesh(v, t)
    a_ref, b_ref = ofem.fem(mesh)
    ta, tb, tl = _triplets(mesh.v, mesh.t, "tria")
    fem = lapy_b200.Solver(mesh)
    _check(fem.stiffness, ta, a_ref, "A")
    _check(fem.mass, tb, b_ref, "B")


def test_full_size_properties():
    """BASELINE.json config 2 size (level-9 icosphere): size-independent properties."""
    import lapy_b200
    from lapy_b200 import mesh as M

    mesh = M.icosphere(9)
    fem = lapy_b200.Solver(mesh)
    a, b = fem.stiffness, fem.mass
    nv, nt = mesh.v.shape[0], mesh.t.shape[0]
    assert a.shape == (nv, nv) and a.nnz == nv + 3 * nt == 18350082
    assert a.has_sorted_indices and np.all(np.diff(a.indptr) >= 6)
    rows = np.repeat(np.arange(nv), np.diff(a.indptr))
    assert np.all(np.diff(a.indices)[rows[1:] == rows[:-1]] > 0)  # strictly increasing per row
    assert np.abs(a @ np.ones(nv)).max() < 1e-9  # constants in the null space
    assert abs(b.sum() - 4 * np.pi) < 1e-4  # area of the unit sphere
    assert (a != a.T).nnz == 0 and (b != b.T).nnz == 0  # bitwise symmetric
    # spot-check 4096 random rows against the oracle's locals for those rows only
    sub = np.unique(np.random.default_rng(0).integers(0, nt, 4096))
    a12, a23, a31, vol = ofem.tria_local(mesh.v, mesh.t[sub])
    t = mesh.t[sub]
    got = np.asarray(a[t[:, 0], t[:, 1]]).ravel()
    # each edge entry = a12 of this triangle + the matching cot of the neighbour: check bound
    assert np.all(np.isfinite(got))
    lm = lapy_b200.Solver(mesh, lump=True).mass
    assert lm.nnz == nv and abs(lm.sum() - b.sum()) < 1e-9


def test_level9_full_csr_vs_oracle():
    """BASELINE.json config 2 (level-9 icosphere, 5,242,880 triangles): the COMPLETE matrices against
    the oracle's SciPy COO->CSC result (~10 s of CPU): structure bit-exact, off-diagonals (two addends:
    order independent) bit-equal, diagonals (six addends) within 1e-12 relative, lumped mass too."""
    import lapy_b200
    from lapy_b200 import mesh as M

    mesh = M.icosphere(9)
    fem = lapy_b200.Solver(mesh)
    a_ref, b_ref = ofem.fem(mesh)
    for got, ref, name in ((fem.stiffness, a_ref, "A"), (fem.mass, b_ref, "B")):
        assert got.shape == ref.shape and got.nnz == ref.nnz == 18350082, name
        np.testing.assert_array_equal(got.indptr, ref.indptr, err_msg=name)
        np.testing.assert_array_equal(got.indices, ref.indices, err_msg=name)
        off = got.indices != np.repeat(np.arange(got.shape[0], dtype=np.int32), np.diff(got.indptr))
        assert np.array_equal(got.data[off], ref.data[off]), name + " off-diagonals bitwise"
        assert np.all(np.abs(got.data - ref.data) <= 1e-12 * np.abs(ref.data)), name + " values 1e-12 relative"
    _, bl_ref = ofem.fem(mesh, lump=True)
    bl = lapy_b200.Solver(mesh, lump=True).mass
    np.testing.assert_array_equal(bl.indices, bl_ref.indices)
    assert np.all(np.abs(bl.data - bl_ref.data) <= 1e-12 * np.abs(bl_ref.data))


def test_cube121_sampled_columns_vs_oracle():
    """BASELINE.json config 3 size (121^3 vertices, 10,368,000 tets): the reference cannot build this
    matrix in reasonable memory/time here, so the COMPLETE columns of ~3000 sampled vertices (corners,
    edges, faces, interior) are rebuilt by the oracle from every tet touching them, in COO input order,
    and compared bit for bit (structure and values); plus global invariants."""
    import lapy_b200
    from lapy_b200 import mesh as M

    n = 121
    mesh = M.cube_tets(n)
    nv = mesh.v.shape[0]
    fem = lapy_b200.Solver(mesh)
    a, b = fem.stiffness, fem.mass
    assert a.shape == (nv, nv) and a.nnz == b.nnz == 26223481  # V + 2E, SURVEY.md §8
    assert np.abs(a @ np.ones(nv)).max() < 1e-9 and abs(b.sum() - 1.0) < 1e-12
    rng = np.random.default_rng(5)
    corners = [x + n * y + n * n * z for x in (0, n - 1) for y in (0, n - 1) for z in (0, n - 1)]
    sample = np.unique(np.concatenate([corners, np.arange(0, n), np.arange(n * n * 60, n * n * 60 + 2 * n),
                                       rng.integers(0, nv, 2500)]))  # fmt: skip
    touched = np.isin(mesh.t, sample).any(axis=1)
    sub = mesh.t[touched]  # ascending element order is kept: same COO input order within every column
    ta, tb, _ = _triplets(mesh.v, sub, "tet")
    for got, trip, name in ((a, ta, "A"), (b, tb, "B")):
        indptr, indices, data = coo_to_csc_sequential(*trip, n=nv)
        for col in sample:
            g0, g1 = got.indptr[col], got.indptr[col + 1]
            r0, r1 = indptr[col], indptr[col + 1]
            assert np.array_equal(got.indices[g0:g1], indices[r0:r1]), (name, col)
            assert np.array_equal(got.data[g0:g1], data[r0:r1]), (name, col)


def test_which_assembly_pipeline_runs():
    """Regular triangle meshes (valence <= 8, no repeated vertex, no degenerate element) must take the
    strip-cooperative kernels - a silent fallback to the record pipeline would be a 3x slowdown nobody
    notices; tets and irregular triangle meshes take the record pipeline."""
    import lapy_b200
    from lapy_b200 import _lib
    from lapy_b200 import mesh as M
    from lapy_b200.mesh import TriaMesh

    ctx = _lib.default_context()
    c0 = ctx.counters()
    lapy_b200.Solver(M.icosphere(5))
    lapy_b200.Solver(M.icosphere(5), lump=True)
    lapy_b200.Solver.fem_tria_mass(M.icosphere(4))
    c1 = ctx.counters()
    assert c1["strip_assemblies"] - c0["strip_assemblies"] == 3 and c1["record_assemblies"] == c0["record_assemblies"]
    lapy_b200.Solver(M.cube_tets(6))
    n = 12
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    fan = TriaMesh(np.vstack([[0, 0, 0.3], np.column_stack([np.cos(ang), np.sin(ang), np.zeros(n)])]),
                   np.column_stack([np.zeros(n, int), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n]))  # fmt: skip
    lapy_b200.Solver(fan)  # valence 12
    c2 = ctx.counters()
    assert c2["record_assemblies"] - c1["record_assemblies"] == 2 and c2["strip_assemblies"] == c1["strip_assemblies"]
