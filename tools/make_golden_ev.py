"""Golden vectors for lapy_b200.io and the ShapeDNA post-processing helpers, made with the UNMODIFIED
reference (dev container only): python tools/make_golden_ev.py -> tests/golden/ev_*.{ev,npz}"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refshim  # noqa: E402

lapy, data = refshim.load()
from lapy import TriaMesh, io, shapedna  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
rng = np.random.default_rng(7)
full = {"Refine": 0, "Degree": 1, "Dimension": 2, "Elements": 20, "DoF": 12, "NumEW": 4, "Area": 12.5, "Volume": 3.25,
        "BLength": 0.0, "EulerChar": 2, "TimePre": 1, "TimeCalcAB": 2, "TimeCalcEW": 3,
        "Eigenvalues": rng.standard_normal(4) * 1e3, "Eigenvectors": rng.standard_normal((12, 4))}  # fmt: skip
io.write_ev(os.path.join(out, "ev_full.ev"), full)
io.write_ev(os.path.join(out, "ev_values_only.ev"), {"NumEW": 5, "Eigenvalues": np.arange(5) * 0.1})
back = io.read_ev(os.path.join(out, "ev_full.ev"))
ico = TriaMesh.read_off(os.path.join(data, "icosahedron.off"))
ico.refine_(2)
ev = np.array([0.0, 2.1, 2.2, 2.3, 6.4, 6.5])
np.savez(
    os.path.join(out, "ev_post.npz"),
    full_eigenvalues=full["Eigenvalues"], full_eigenvectors=full["Eigenvectors"],
    back_keys=np.array(sorted(back.keys())), back_eigenvalues=back["Eigenvalues"], back_eigenvectors=back["Eigenvectors"],
    cube_tria_ev=io.read_ev(os.path.join(data, "cubeTria.ev"))["Eigenvalues"],
    ico_v=ico.v, ico_t=ico.t, ico_area=ico.area(), ev=ev,
    norm_surface=shapedna.normalize_ev(ico, ev, method="surface"), norm_geometry=shapedna.normalize_ev(ico, ev),
    reweighted=shapedna.reweight_ev(ev), distance=shapedna.compute_distance(ev, ev[::-1].copy()),
)  # fmt: skip
print("written", sorted(f for f in os.listdir(out) if f.startswith("ev_")))
