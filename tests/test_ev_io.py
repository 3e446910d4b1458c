"""ShapeDNA .ev files and post-processing against golden files written by the
unmodified reference (tools/make_golden_ev.py).  Host-only code: runs without a GPU and without the
CUDA library."""

import filecmp
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _full():
    g = np.load(os.path.join(GOLD, "ev_post.npz"))
    return {"Refine": 0, "Degree": 1, "Dimension": 2, "Elements": 20, "DoF": 12, "NumEW": 4, "Area": 12.5, "Volume": 3.25,
            "BLength": 0.0, "EulerChar": 2, "TimePre": 1, "TimeCalcAB": 2, "TimeCalcEW": 3,
            "Eigenvalues": g["full_eigenvalues"], "Eigenvectors": g["full_eigenvectors"]}, g  # fmt: skip


def test_write_ev_is_byte_identical_to_the_reference(tmp_path):
    from lapy_b200.io import write_ev

    d, _ = _full()
    write_ev(str(tmp_path / "a.ev"), d)
    assert filecmp.cmp(tmp_path / "a.ev", os.path.join(GOLD, "ev_full.ev"), shallow=False)
    write_ev(str(tmp_path / "b.ev"), {"NumEW": 5, "Eigenvalues": np.arange(5) * 0.1})
    assert filecmp.cmp(tmp_path / "b.ev", os.path.join(GOLD, "ev_values_only.ev"), shallow=False)
    with pytest.raises(ValueError, match="no Eigenvalues"):
        write_ev(str(tmp_path / "c.ev"), {"NumEW": 5})


def test_read_ev_matches_the_reference():
    from lapy_b200.io import read_ev

    d, g = _full()
    back = read_ev(os.path.join(GOLD, "ev_full.ev"))
    assert sorted(back.keys()) == list(g["back_keys"])
    np.testing.assert_array_equal(back["Eigenvalues"], g["back_eigenvalues"])
    np.testing.assert_array_equal(back["Eigenvectors"], g["back_eigenvectors"])
    np.testing.assert_array_equal(back["EigenvectorsSize"], [12, 4])
    for k in ("Refine", "Degree", "Dimension", "Elements", "DoF", "NumEW", "EulerChar"):
        assert back[k] == d[k] and isinstance(back[k], int)
    for k in ("Area", "Volume", "BLength"):
        assert back[k] == d[k] and isinstance(back[k], float)
    only = read_ev(os.path.join(GOLD, "ev_values_only.ev"))
    assert set(only) == {"NumEW", "Eigenvalues"}
    np.testing.assert_array_equal(only["Eigenvalues"], np.arange(5) * 0.1)
    with pytest.raises(OSError):
        read_ev(os.path.join(GOLD, "does_not_exist.ev"))


def test_round_trip_with_text_header_and_single_column(tmp_path):
    """Two files the reference writes but cannot read back (int('name'), '(…)' on the brace line)."""
    from lapy_b200.io import read_ev, write_ev

    d = {"Creator": "lapy_b200", "User": "someone", "File": "mesh.vtk", "NumEW": 2, "Eigenvalues": np.array([0.0, 1.5]),
         "Eigenvectors": np.array([[1.0], [2.0], [3.0]])}  # fmt: skip
    write_ev(str(tmp_path / "t.ev"), d)
    back = read_ev(str(tmp_path / "t.ev"))
    assert back["Creator"] == "lapy_b200" and back["User"] == "someone" and back["File"] == "mesh.vtk"
    np.testing.assert_array_equal(back["Eigenvalues"], d["Eigenvalues"])
    np.testing.assert_array_equal(back["Eigenvectors"], d["Eigenvectors"])


def test_shapedna_post_processing_matches_the_reference():
    from lapy_b200 import shapedna
    from lapy_b200.mesh import TetMesh, TriaMesh

    _, g = _full()
    ico = TriaMesh(g["ico_v"], g["ico_t"])
    assert ico.area() == g["ico_area"]
    ev = g["ev"]
    np.testing.assert_array_equal(shapedna.normalize_ev(ico, ev, method="surface"), g["norm_surface"])
    np.testing.assert_array_equal(shapedna.normalize_ev(ico, ev), g["norm_geometry"])
    np.testing.assert_array_equal(shapedna.reweight_ev(ev), g["reweighted"])
    assert shapedna.compute_distance(ev, ev[::-1].copy()) == pytest.approx(float(g["distance"]), rel=1e-15)
    with pytest.raises(ValueError, match="not implemented"):  # lapy/shapedna.py:292-296
        shapedna.compute_distance(ev, ev, dist="other")
    with pytest.raises(ValueError, match="Unknown normalization"):
        shapedna.normalize_ev(ico, ev, method="nope")
    # enclosed volume of the closed, oriented surface (lapy/shapedna.py:50-93, lapy/tria_mesh.py:671-702)
    np.testing.assert_allclose(shapedna.normalize_ev(ico, ev, method="volume"), ev * ico.volume() ** (2.0 / 3.0), rtol=1e-15)
    tet = TetMesh(np.eye(4, 3), np.array([[0, 1, 2, 3]]))
    with pytest.raises(NotImplementedError):
        shapedna.normalize_ev(tet, ev)
    flat = TriaMesh(np.zeros((3, 3)), np.array([[0, 1, 2]]))
    with pytest.raises(ValueError, match="positive"):
        shapedna.normalize_ev(flat, ev, method="surface")


def test_heat_kernel_and_diagonal():
    """heat.diagonal equals the reference's expression (lapy/heat.py:58) on its documented shapes
    (column of eigenvalues, row of times); heat.kernel follows the docstring formula."""
    from lapy_b200 import heat

    rng = np.random.default_rng(1)
    evecs = rng.standard_normal((30, 6))
    evals = np.abs(rng.standard_normal((6, 1))).cumsum(0)
    t = np.array([[0.1, 0.5, 2.0]])
    x = np.array([0, 3, 7])
    d = heat.diagonal(t, x, evecs, evals, 4)
    ref = np.matmul(evecs[x, 0:4] * evecs[x, 0:4], np.exp(-np.matmul(evals[0:4], t)))
    np.testing.assert_array_equal(d, ref)
    k = heat.kernel(t, 5, evecs, evals, 4)
    assert k.shape == (30, 3)
    want = sum(np.exp(-evals[j, 0] * t[0])[None, :] * evecs[:, j : j + 1] * evecs[5, j] for j in range(4))
    np.testing.assert_allclose(k, want, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(heat.kernel(0.5, 5, evecs, evals[:, 0], 4)[:, 0], k[:, 1], rtol=1e-13)
    np.testing.assert_allclose(np.diag(np.column_stack([heat.kernel(0.5, v, evecs, evals, 4)[:, 0] for v in range(30)])),
                               heat.diagonal(np.array([[0.5]]), np.arange(30), evecs, evals, 4)[:, 0], rtol=1e-13)  # fmt: skip
    for f in (heat.diagonal, heat.kernel):
        with pytest.raises(ValueError, match="exceeds"):
            f(t, 0 if f is heat.kernel else x, evecs, evals, 7)
