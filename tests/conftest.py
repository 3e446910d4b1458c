import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_mesh(g):
    from lapy_b200.mesh import TetMesh, TriaMesh

    return (TetMesh if g["t"].shape[1] == 4 else TriaMesh)(g["v"], g["t"])


def golden_csc(g, name):
    from scipy import sparse

    return sparse.csc_matrix(
        (g[name + "_data"], g[name + "_indices"], g[name + "_indptr"]), shape=tuple(g[name + "_shape"])
    )


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get
