"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports every
symbol include/lapy_b200.h declares, the ctypes table binds each of them, and the product never
imports the oracle."""

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lapy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g

    g.build()
    from lapy_b200 import _lib

    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/lapy_b200.h but not exported: {missing}"
    unbound = [n for n in names if n not in _lib.SIGNATURES and n not in _lib.STRING_GETTERS]
    assert not unbound, f"exported but not bound in lapy_b200/_lib.py: {unbound}"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (lb_[a-z0-9_]+)", out)))
    undeclared = [n for n in exported if n not in names]
    assert not undeclared, f"exported without a declaration in the header: {undeclared}"


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lapy_b200 import _lib

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lapy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "scipy.sparse.linalg" not in txt, f"{f}: product must not call SciPy solvers"


def test_sm100a_only():
    out = subprocess.run(
        ["cuobjdump", "--list-elf", os.path.join(pkg := os.path.join(ROOT, "lapy_b200"), "liblapyb200.so")],
        capture_output=True, text=True,
    ).stdout  # fmt: skip
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_result_pool_falls_back_without_a_gpu():
    """Large result arrays come from a pool of page-locked blocks (lb_host_alloc); without a usable CUDA device
    (this suite) or below the size threshold the pool hands out ordinary NumPy arrays of the requested shape."""
    import numpy as np

    from lapy_b200 import _lib

    small = _lib._pinned.empty((10, 3))
    assert small.shape == (10, 3) and small.dtype == np.float64 and small.flags.c_contiguous
    big = _lib._pinned.empty((4_500_000, 1))  # 36 MB: above the threshold
    assert big.shape == (4_500_000, 1) and big.dtype == np.float64 and big.flags.writeable
    big[:] = 1.0
    assert float(big.sum()) == 4_500_000.0
