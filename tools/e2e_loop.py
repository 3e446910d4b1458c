"""Dev tool: wall time of consecutive compute_shapedna calls from host arrays (the e2e step of bench.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lapy_b200 import _lib, mesh as M
from lapy_b200.shapedna import compute_shapedna

if os.environ.get("POOL", "1") == "0":
    _lib._pinned.max_bytes = 0
mesh = M.icosphere(9)
sd = None
for i in range(7):
    t0 = time.perf_counter()
    sd = compute_shapedna(mesh, k=50)
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
