"""CPU oracle for the LaPy FEM hot path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy/SciPy restatement of the reference's algorithm for the path in SURVEY.md §8a
(Deep-MI/LaPy v1.6.0-dev: lapy/solver.py, lapy/heat.py, lapy/diffgeo.py, lapy/shapedna.py).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed CPU arm -
never as something the product path falls back to.  ``lapy_b200`` never imports ``oracle``.

Parity is PINNED: ``tools/make_golden.py`` ran the unmodified reference in the build
container and stored its inputs/outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every oracle function against them (assembly bit-exact, solves to 1e-9) and against the
golden vectors the reference's own tests hold (data/cubeTria.ev, data/cubeTetra.ev,
lapy/utils/tests/expected_outcomes.json values quoted in the tests).

The third-party arithmetic on the path is *called*, not restated, because
the same binaries exist here and on the GPU box: SciPy (reference pins ``scipy!=1.13.0``,
pyproject.toml:50; this image: 1.18.x) for COO->CSC (``coo_tocsr`` + ``csr_sort_indices`` +
``csr_sum_duplicates``), SuperLU ``splu`` and ARPACK ``eigsh``.  ``oracle.sparse_build`` also
restates the COO->CSC algorithm in pure NumPy to document the summation order the CUDA kernels
use.
"""

from . import diffgeo, fem, solve, sparse_build  # noqa: F401
