#!/usr/bin/env python
"""bench.py - the driver's benchmark contract for lapy_b200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

N=1 workload = BASELINE.json configs[1]: synthetic level-9 icosphere (2,621,442 vertices,
5,242,880 triangles), one "step" = one pass of the hot path over one mesh.  With N>1 (torchrun,
one rank per GPU) every rank processes its own mesh per step (mesh-parallel batch, no data-path
collective) -> "scaling": "weak"; value = meshes all ranks processed / max-over-ranks device time.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU implementation of
the same path (the oracle: NumPy element math + SciPy COO->CSC / SuperLU / ARPACK) on the host.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PEAK_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return PEAK_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()  # fmt: skip
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(sm),
        }


def make_workload(name, rank):
    from lapy_b200 import mesh as M

    if name == "icosphere9":
        return M.icosphere(9), "level-9 icosphere, 2,621,442 v / 5,242,880 tris (BASELINE.json configs[1])"
    if name.startswith("icosphere"):
        lvl = int(name[len("icosphere"):])
        return M.icosphere(lvl), f"level-{lvl} icosphere (reduced size: NOT the headline config)"
    if name.startswith("cube"):
        n = int(name[len("cube"):])
        return M.cube_tets(n), f"structured tet cube n={n}"
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(mesh, nnz):
    """read t and v once, write A and B once (fp64 values, int32 indices)."""
    nt, k = mesh.t.shape
    nv = mesh.v.shape[0]
    return 4 * k * nt + 24 * nv + 2 * (12 * nnz + 4 * (nv + 1))


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm (oracle) on the host cores, bounded sample."""
    if rank != 0:
        return
    from oracle import fem as ofem

    mesh, desc = make_workload(args.workload, 0)
    nt = mesh.t.shape[0]
    for _ in range(max(args.warmup, 1) if args.steps > 1 else 1):
        ofem.fem(mesh)
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        a, b = ofem.fem(mesh)
    dt = (time.perf_counter() - t0) / steps
    val = nt / dt / 1e9
    line = {
        "impl": "reference", "metric": "fem_assembly_gelem_per_s", "value": val, "unit": "Gelem/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "step": "Solver(mesh): stiffness + full mass assembly to CSC"},
        "cpu_baseline": {"value": val, "unit": "Gelem/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full assemblies of the same mesh (NumPy + SciPy COO->CSC, sequential)"},
        "e2e": {"value": val, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="icosphere9")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import lapy_b200
    from lapy_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)

    mesh, desc = make_workload(args.workload, rank)
    nt = mesh.t.shape[0]
    kind = _lib.FEM_TETRA if mesh.t.shape[1] == 4 else _lib.FEM_TRIA
    dmesh = _lib.DeviceMesh(ctx, mesh.v, mesh.t)  # inputs resident in HBM before the timed region

    def step():
        dmesh.drop_cache()  # the vertex->element incidence is part of the assembly work
        a, b = _lib.assemble(ctx, dmesh, kind, False)
        return a, b

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        a, b = step()
    nnz = a.nnz
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local_rank) as clk:
        ctx.timer_start()
        for _ in range(args.steps):
            a, b = step()
        ms = ctx.timer_stop()
    launches = ctx.launch_count() - l0
    barrier()
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps

    # ---- end to end through the public API with host buffers (pinned), H2D + D2H inside ----
    vpin = torch.from_numpy(np.ascontiguousarray(mesh.v)).pin_memory().numpy()
    tpin = torch.from_numpy(np.ascontiguousarray(mesh.t)).pin_memory().numpy()
    hmesh = type(mesh)(vpin, tpin)
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        hmesh.__dict__.pop("_lb_device_mesh", None)
        fem = lapy_b200.Solver(hmesh, ctx=ctx)
        return fem.stiffness, fem.mass

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sa, sb = e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h2d = vpin.nbytes + tpin.nbytes
    d2h = 2 * (sa.data.nbytes + sa.indices.nbytes + sa.indptr.nbytes)

    if rank == 0:
        peak, peak_kind = measured_peak()
        abytes = algorithmic_bytes(mesh, nnz)
        achieved = abytes / (ms_step * 1e-3) / 1e9
        line = {
            "metric": "fem_assembly_gelem_per_s", "value": world * nt / (ms_step * 1e-3) / 1e9, "unit": "Gelem/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "step": "Solver(mesh): stiffness + full mass assembly to CSC",
                       "l2": "inputs+outputs per step (629 MB) exceed the 126 MB L2", "parallelism": f"mesh-parallel x{world}"},
            "gpu_launches": int(launches),
            "e2e": {"value": world * nt / e2e_s / 1e9, "unit": "Gelem/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": peak_kind, "kernel": "assembly pipeline (all kernels of one step)",
                         "algorithmic_bytes": int(abytes)},
            "clocks": clk.summary(),
        }  # fmt: skip
        if not args.no_cpu_baseline:
            from oracle import fem as ofem

            ofem.fem(mesh)
            t0 = time.perf_counter()
            ofem.fem(mesh)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nt / dt / 1e9, "unit": "Gelem/s", "cores": 1, "kind": "port",
                                    "sample": "1 full assembly of the same mesh (NumPy + SciPy COO->CSC, sequential)"}  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
