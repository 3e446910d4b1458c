#!/usr/bin/env python
"""bench.py - the driver's benchmark contract for lapy_b200 (DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--k 50]

N=1 workload = BASELINE.json configs[1]: ShapeDNA k=50 of the synthetic level-9 icosphere
(2,621,442 vertices / 5,242,880 triangles).  One "step" = one pass of the hot path over one mesh:
FEM assembly of stiffness + mass (incl. the vertex->element incidence) and the generalized
eigensolve, from a mesh already resident in HBM to eigenvalues / eigenvectors in host NumPy arrays.
With N>1 (torchrun, one rank per GPU) every rank processes its own mesh per step (mesh-parallel
batch as in BrainPrint, no data-path collective) -> "scaling": "weak"; value = meshes all ranks
processed / max-over-ranks device time.

The eigenvalues of the LAST TIMED step are compared with the reference's (tests/golden/spectra.npz,
produced by the unmodified reference); a relative error above 1e-8 makes the run fail (rc 3).

Prints ONE JSON line (rank 0).  Sub-records (outside the timed region, N=1 unless noted):
`configs` = BASELINE.json configs 3, 4, 5 at full size (tet cube 121^3, heat + geodesics on level 9,
a batch of level-7 surfaces; the batch also at N>1), `rowpart` (N>1) = one mesh solved cooperatively
by all ranks (row-partitioned, NCCL halo exchange).

`--impl reference` times the reference's CPU algorithm for the same path (the oracle: NumPy element
math + SciPy COO->CSC + SuperLU + ARPACK, sequential per mesh) for real on the largest icosphere
level that fits a few minutes (level 8), one mesh per rank's worth of host processes.
"""

from __future__ import annotations

import argparse
import faulthandler
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
faulthandler.enable()  # a crash inside the native library prints the Python stack on stderr

PEAK_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
METRIC = "shapedna_k50_meshes_per_s"
PARITY_RTOL = 1e-8  # BASELINE.json north_star
ROWPART_WORLDS = {2, 4, 8}  # world sizes the row-partitioned solve has been validated on (profiles/rowpart_*_r2.log)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return PEAK_FALLBACK_GBS, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML during the timed region."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=3)

    def summary(self):
        return {
            "sm_mhz": float(np.median(self.sm)) if self.sm else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.sm),
        }


def make_workload(name, rank=0):
    """-> (mesh, description, key of the reference spectrum in tests/golden/spectra.npz or None)"""
    from lapy_b200 import mesh as M

    if name == "icosphere9":
        return (M.icosphere(9), "level-9 icosphere, 2,621,442 v / 5,242,880 tris, ShapeDNA k=50 (BASELINE.json configs[1])",
                "ico9_k50")  # fmt: skip
    if name.startswith("icosphere"):
        lvl = int(name[len("icosphere"):])
        return M.icosphere(lvl), f"level-{lvl} icosphere (reduced size: NOT the headline config)", f"ico{lvl}_k50"
    if name.startswith("brain"):  # BrainPrint-like batch surface (config 5)
        return M.perturbed_sphere(7, seed=rank), "level-7 perturbed sphere, 163,842 v (BASELINE.json configs[4] unit)", None
    if name.startswith("cube"):
        n = int(name[len("cube"):])
        return M.cube_tets(n), f"structured tet cube n={n}", f"cube{n}_k50"
    raise SystemExit(f"unknown workload {name}")


def golden_spectrum(key, k):
    if key is None or k != 50:
        return None
    p = os.path.join(ROOT, "tests", "golden", "spectra.npz")
    if not os.path.exists(p):
        return None
    g = np.load(p)
    return np.asarray(g[key]) if key in g else None


def parity_record(ev, ref, key):
    """Eigenvalues of the timed step against the unmodified reference's (relative, lambda_0 absolute)."""
    if ref is None:
        return {"checked": False, "why": "no reference spectrum for this workload / k in tests/golden/spectra.npz"}
    ev, ref = np.asarray(ev), np.asarray(ref)
    rel = float(np.max(np.abs(ev[1:] - ref[1:]) / np.abs(ref[1:])))
    ok = bool(rel <= PARITY_RTOL and abs(ev[0]) <= 1e-8 and np.all(np.isfinite(ev)))
    return {"checked": True, "ok": ok, "max_rel_err": rel, "rtol": PARITY_RTOL, "lam0": float(ev[0]), "lam0_ref": float(ref[0]),
            "golden": f"tests/golden/spectra.npz[{key}] (unmodified reference, BASELINE.md §5.2)"}  # fmt: skip


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (= the reference's algorithm on SciPy SuperLU + ARPACK)
# ------------------------------------------------------------------------------------------------
def _cpu_shapedna_worker(level, k, q):
    from lapy_b200 import mesh as M
    from oracle import solve as osolve

    mesh = M.icosphere(level)
    t0 = time.perf_counter()
    sd = osolve.shapedna(mesh, k=k)
    q.put((time.perf_counter() - t0, float(sd["Eigenvalues"][1])))


def cpu_shapedna_seconds(level, k, procs=1):
    """Wall seconds for `procs` concurrent reference ShapeDNA runs (one sequential process each) of
    the level-`level` icosphere; mesh generation is outside the timing."""
    import multiprocessing as mp

    if procs == 1:
        from lapy_b200 import mesh as M
        from oracle import solve as osolve

        mesh = M.icosphere(level)
        t0 = time.perf_counter()
        osolve.shapedna(mesh, k=k)
        return time.perf_counter() - t0, [time.perf_counter() - t0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    # the BLAS inside SuperLU / ARPACK is threaded: split the host threads over the processes
    # (oversubscription makes concurrent runs several times slower than sequential ones)
    threads = str(max(1, (os.cpu_count() or 1) // procs))
    saved = {v: os.environ.get(v) for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for v in saved:
        os.environ[v] = threads
    try:
        ps = [ctx.Process(target=_cpu_shapedna_worker, args=(level, k, q)) for _ in range(procs)]
        for p in ps:
            p.start()
    finally:
        for v, old in saved.items():
            if old is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = old
    each = [q.get()[0] for _ in ps]
    for p in ps:
        p.join()
    return max(each), each


NV = {4: 2562, 5: 10242, 6: 40962, 7: 163842, 8: 655362, 9: 2621442}


def run_reference(args, rank):
    """Reference arm: a REAL measurement of the largest level that fits a few minutes (level 8,
    ~2-4.5 min), labelled as executed.  The level-9 workload costs 13-26 min and 27 GB per mesh on
    the CPU (1557 s measured offline), so it is only PROJECTED, in a separate field,
    with the growth exponent measured in this run (level 7 -> 8)."""
    if rank != 0:
        return
    n_units = max(1, args.gpus)  # one mesh per GPU of the arm it is compared with
    cores = os.cpu_count() or 1
    procs = min(n_units, cores)
    level = args.ref_level
    try:
        import psutil

        if psutil.virtual_memory().available < procs * (9 << 30) and level >= 8:  # SuperLU at level 8: ~7 GB per process
            level = 7
    except Exception:
        pass
    t_small, _ = cpu_shapedna_seconds(level - 1, args.k, 1)  # also the warm-up (imports, page cache)
    rounds = (n_units + procs - 1) // procs
    wall = 0.0
    each = []
    for _ in range(rounds):
        w, e = cpu_shapedna_seconds(level, args.k, procs)
        wall += w
        each += e
    exponent = float(np.log(np.mean(each) / t_small) / np.log(NV[level] / NV[level - 1]))
    proj9 = float(np.mean(each) * (NV[9] / NV[level]) ** exponent)
    val = n_units / wall
    what = (f"{n_units} x ShapeDNA k={args.k} of the level-{level} icosphere ({NV[level]:,} v) run for real by the oracle "
            f"(SciPy SuperLU + ARPACK: sequential algorithms, threaded BLAS), {procs} concurrent process(es) on {cores} host threads: {wall:.1f} s wall; "
            f"level {level - 1} took {t_small:.1f} s -> time ~ n^{exponent:.2f}")  # fmt: skip
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "meshes/s", "n_gpus": args.gpus, "steps": 1,
        "warmup": 0, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"level-{level} icosphere, {NV[level]:,} v, ShapeDNA k={args.k}: the LARGEST level the CPU reference "
                               "finishes within a few minutes - NOT the level-9 workload of the b200 arm (4x the vertices)",
                   "step": "Solver(mesh) + eigs(k): assembly + splu + ARPACK shift-invert", "k": args.k,
                   "meshes_per_step": n_units},
        "cpu_baseline": {"value": val, "unit": "meshes/s", "cores": cores if procs == 1 else procs * max(1, cores // procs),
                         "kind": "port", "sample": what},
        "e2e": {"value": val, "unit": "meshes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "level9_projection": {"s_per_mesh": proj9, "meshes_per_s": n_units / proj9 / rounds if rounds else None,
                              "how": f"level-{level} time x 4^{exponent:.2f} (exponent measured in this run); "
                                     "measured offline on the survey box: 1557 s/mesh (BASELINE.md §2)"},
        "host": {"cpu_count": cores},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(args):
    """cpu_baseline of the b200 arm: ~15-30 s of CPU work = the oracle's ShapeDNA on ONE level-7
    icosphere (1/16 of the workload's vertices), reported as measured - not scaled."""
    t, _ = cpu_shapedna_seconds(7, args.k, 1)
    return {"value": 1.0 / t, "unit": "meshes/s", "cores": 1, "kind": "port",
            "sample": f"1 ShapeDNA k={args.k} of a level-7 icosphere (163,842 v, 1/16 of the workload mesh) = {t:.1f} s on one host "
                      "core; value is for THAT mesh size, not scaled. The level-9 mesh itself: 1557 s measured offline "
                      "(BASELINE.md §2); `bench.py --impl reference` runs level 8 for real",
            "sample_seconds": t, "sample_vertices": NV[7]}  # fmt: skip


# ------------------------------------------------------------------------------------------------
# sub-records: BASELINE.json configs 3, 4, 5 (outside the timed region)
# ------------------------------------------------------------------------------------------------
def extras_single_gpu(ctx, peak, workload_mesh=None):
    import lapy_b200
    from lapy_b200 import _lib, diffgeo, heat
    from lapy_b200 import mesh as M

    out = {}

    def timed(f):
        ctx.sync()
        t0 = time.perf_counter()
        r = f()
        ctx.sync()
        return r, time.perf_counter() - t0

    # config 3 on one GPU: 121^3 tet cube (1,771,561 v / 10,368,000 tets): assembly + eigs(k=50)
    try:
        mesh = M.cube_tets(121)
        dm = _lib.DeviceMesh(ctx, mesh.v, mesh.t)
        ms = []
        for _ in range(4):
            dm.drop_cache()
            ctx.timer_start()
            a, b = _lib.assemble(ctx, dm, _lib.FEM_TETRA, False)
            ms.append(ctx.timer_stop())
        nt, nv, nnz = mesh.t.shape[0], mesh.v.shape[0], a.nnz
        algo = 16 * nt + 24 * nv + 2 * (12 * nnz + 4 * (nv + 1))
        _lib.eigs(ctx, a, b, 50, -0.01)
        (ev, _evec, info), t_eig = timed(lambda: _lib.eigs(ctx, a, b, 50, -0.01))
        asm = float(np.median(ms[1:]))
        out["cube121_tets"] = {"vertices": nv, "tets": nt, "nnz": int(nnz), "assembly_ms": asm, "assembly_gelem_per_s": nt / asm / 1e6,
                               "assembly_roofline_frac": algo / (asm * 1e-3) / 1e9 / peak, "eigs_k50_s": t_eig,
                               "iterations": info["iterations"], "residual": info["residual"],
                               "lam1_over_pi2": float(ev[1] / np.pi**2), "lam0": float(ev[0])}  # fmt: skip
        del a, b, dm, _evec
    except Exception as e:  # a sub-record must not take the headline down
        out["cube121_tets"] = {"error": repr(e)}
    # config 4: heat diffusion + heat-method geodesics on the level-9 icosphere
    try:
        mesh = workload_mesh if workload_mesh is not None and workload_mesh.v.shape[0] == 2621442 else M.icosphere(9)
        heat.diffusion(mesh, [0], m=1.0)
        u, t_heat = timed(lambda: heat.diffusion(mesh, [0], m=1.0))
        hinfo = dict(heat.diffusion.last_info)
        # m=1 (t = h^2 = 5.6e-6) underflows beyond ~1.7 rad for the reference as well; m=16 gives the
        # diffusion length of level 7 with m=1 and a usable geodesic field
        u16, t_heat16 = timed(lambda: heat.diffusion(mesh, [0], m=16.0))
        hinfo16 = dict(heat.diffusion.last_info)
        diffgeo.compute_geodesic_f(mesh, u16)
        g, t_geo = timed(lambda: diffgeo.compute_geodesic_f(mesh, u16))
        exact = np.arccos(np.clip(mesh.v @ mesh.v[0], -1, 1))
        out["heat_geodesic_L9"] = {"heat_m1_s": t_heat, "heat_m1_sweeps": hinfo["iterations"], "heat_m16_s": t_heat16,
                                   "heat_m16_sweeps": hinfo16["iterations"], "u0_m1": float(u[0]), "geodesic_s": t_geo,
                                   "geodesic_max": float(g.max()), "max_abs_err_vs_great_circle": float(np.abs(g - exact).max())}  # fmt: skip
    except Exception as e:
        out["heat_geodesic_L9"] = {"error": repr(e)}
    return out


def batch_record(world, workers=4, per_rank=12):
    """config 5 (scaled down to `per_rank` surfaces per GPU): level-7 perturbed spheres, ShapeDNA k=50
    each through lapy_b200.batch.batched_shapedna (round-robin over the ranks, `workers` concurrent
    contexts per GPU)."""
    import torch.distributed as dist

    from lapy_b200 import mesh as M
    from lapy_b200.batch import batched_shapedna

    rank = dist.get_rank() if world > 1 else 0
    total = per_rank * world
    cache = {i: M.perturbed_sphere(7, seed=i) for i in range(rank, total, world)}
    for m in cache.values():
        m.v.flags.writeable = False
        m.t.flags.writeable = False
    batched_shapedna(cache.__getitem__, n_meshes=min(total, max(2, workers) * world), k=50, workers=workers)  # warm-up: every worker context once
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ev = batched_shapedna(cache.__getitem__, n_meshes=total, k=50, workers=workers)
    dt = time.perf_counter() - t0
    return {"meshes": total, "vertices_per_mesh": 163842, "k": 50, "workers_per_gpu": workers, "seconds": dt,
            "meshes_per_s": total / dt, "lam1_range": [float(ev[:, 1].min()), float(ev[:, 1].max())],
            "note": "BASELINE.json configs[4] unit (512 surfaces) scaled to %d per GPU; wall clock incl. H2D / D2H" % per_rank}  # fmt: skip


def rowpart_record(local_rank, world, n_cube=121, k=50):
    """N>1: ONE mesh (BASELINE.json configs[2]: the 121^3 tet cube, 1,771,561 v / 10,368,000 tets) solved
    cooperatively by all ranks: contiguous row blocks of the locality-numbered operator per rank, NCCL
    halo exchange in every SpMM, all-reduced Gram matrices, hierarchy applied column-parallel.  Seconds
    are the max over ranks of the wall time of the second call (the first pays NCCL / cuSOLVER set-up)."""
    import torch
    import torch.distributed as dist

    import lapy_b200
    from lapy_b200 import _lib
    from lapy_b200 import mesh as M

    ctx = _lib.Context(local_rank)
    ctx.init_row_partition()
    mesh = M.cube_tets(n_cube)
    mesh.v.flags.writeable = False
    mesh.t.flags.writeable = False
    fem = lapy_b200.Solver(mesh, ctx=ctx)
    fem.eigs(k=k)
    dist.barrier()
    t0 = time.perf_counter()
    ev, evec = fem.eigs(k=k)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    info = dict(fem.last_info)
    ctx.leave_row_partition()
    n = mesh.v.shape[0]
    return {"workload": f"{n_cube}^3 tet cube, {n:,} v / {mesh.t.shape[0]:,} tets, eigs(k={k}) row-partitioned over {world} GPUs",
            "seconds": float(dt.item()), "lobpcg_device_ms": info["solve_ms"], "iterations": info["iterations"],
            "residual": info["residual"], "lam1_over_pi2": float(ev[1] / np.pi**2), "lam0": float(ev[0]),
            "rows_per_rank": (n + world - 1) // world,
            "note": "every rank receives the full (n, k) eigenvector array (708 MB D2H per rank, inside `seconds`)"}  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="icosphere9")
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--ref-level", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from lapy_b200 import _lib
    from lapy_b200.shapedna import compute_shapedna

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout
        # clean (ONE JSON line) by pointing fd 1 at stderr until the first collective is done
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    ctx = _lib.Context(local_rank)

    mesh, desc, gkey = make_workload(args.workload, rank)
    nt, nv = mesh.t.shape[0], mesh.v.shape[0]
    kind = _lib.FEM_TETRA if mesh.t.shape[1] == 4 else _lib.FEM_TRIA
    dmesh = _lib.DeviceMesh(ctx, mesh.v, mesh.t)  # inputs resident in HBM before the timed region
    asm_ms = []

    # host result buffer of the throughput loop, allocated (and touched) once: a fresh 1 GB NumPy array per
    # step costs 0.1-0.8 s of page faults / munmap on the host, box dependent (measured as step-to-step
    # spread 1.80 .. 2.64 s); the end-to-end number below goes through the public API and allocates per call
    # ... and page-locked: the library then downloads with one DMA instead of staging through its own pinned buffers
    evec_buf = torch.zeros((mesh.v.shape[0], args.k), dtype=torch.float64).pin_memory().numpy()

    def step():
        dmesh.drop_cache()  # the vertex->element incidence is part of the assembly work
        a, b = _lib.assemble(ctx, dmesh, kind, False)
        ev, evec, info = _lib.eigs(ctx, a, b, args.k, -0.01, out_evecs=evec_buf)
        return ev, evec, info, a.nnz

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        ev, evec, info, nnz = step()
    import gc

    gc.collect()  # a generation-2 collection costs ~0.75 s with torch imported: not inside a timed step (see e2e below)
    gc.disable()
    barrier()
    l0 = ctx.launch_count()
    ctx.profile_enable(2)  # CUDA-event pairs around the SpMM launches only (the kernel the roofline is quoted on)
    with ClockSampler(local_rank) as clk:
        ctx.timer_start()  # CUDA events on the library's stream (the stream every kernel runs on)
        step_wall = []
        for _ in range(args.steps):
            t_s = time.perf_counter()
            ev, evec, info, nnz = step()
            step_wall.append((time.perf_counter() - t_s) * 1e3)
        dev_ms = ctx.timer_stop()
    gc.enable()
    print(f"[bench] rank {rank}: per-step wall ms {[round(x, 1) for x in step_wall]}", file=sys.stderr)
    spmm_prof = ctx.profile_report().get("spmm", {"launches": 0, "ms": 0.0, "work": 0.0})
    spmm_shapes = ctx.profile_shapes("spmm", 24)
    ctx.profile_enable(False)
    launches = ctx.launch_count() - l0
    parity = parity_record(ev, golden_spectrum(gkey, args.k), gkey)  # the LAST TIMED step's eigenvalues
    barrier()
    # per-class breakdown from ONE extra step with event pairs around every hot launch (outside the
    # timed region: ~30k event records per step cost device and host time)
    ctx.profile_enable(1)
    ctx.timer_start()
    step()
    prof_step_ms = ctx.timer_stop()
    prof = ctx.profile_report()
    dense_shapes = {cls: ctx.profile_shapes(cls, 8) for cls in ("gram", "update", "small_dense")}
    ctx.profile_enable(False)
    # assembly alone (device time), outside the timed region
    for _ in range(5):
        dmesh.drop_cache()
        ctx.timer_start()
        _lib.assemble(ctx, dmesh, kind, False)
        asm_ms.append(ctx.timer_stop())
    # SpMV / SpMM kernel alone (x, y resident): the north star names SpMV >= 60 % of the HBM roofline
    a_dev, _b = _lib.assemble(ctx, dmesh, kind, False)
    spmv_ms = _lib.spmm_benchmark(ctx, a_dev, 1, 50)
    spmv_ren_ms = _lib.spmm_benchmark(ctx, a_dev, 1, 50, renumber=True)
    spmm64_ren_ms = _lib.spmm_benchmark(ctx, a_dev, 64, 20, renumber=True)
    # the instantiation that took the most device time inside the timed region (level-0 operators only)
    # shape[0] = columns (+ 100000 for the single-precision launches of the multigrid cycle), shape[1] = nnz
    for s in spmm_shapes:
        s["f32"] = s["shape"][0] >= 100000
        s["cols"] = s["shape"][0] % 100000
    lvl0 = [s for s in spmm_shapes if s["shape"][1] == nnz]
    dom = max(lvl0, key=lambda s: s["ms"]) if lvl0 else None
    dom_iso_ms = _lib.spmm_benchmark(ctx, a_dev, dom["cols"], 20, renumber=True, variant=2 if dom["f32"] else 0) if dom else None
    del _b
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps

    # ---- end to end through the public API with host buffers (pinned), H2D + D2H inside ----
    vpin = torch.from_numpy(np.ascontiguousarray(mesh.v)).pin_memory().numpy()
    tpin = torch.from_numpy(np.ascontiguousarray(mesh.t)).pin_memory().numpy()
    hmesh = type(mesh)(vpin, tpin)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        return compute_shapedna(hmesh, k=args.k)  # writable arrays: uploads v / t on every call

    for _ in range(3):  # the binding's pool of page-locked result blocks reaches its steady state (two blocks
        sd = e2e_step()  # alternate while `sd` is rebound) - like the W warm-up steps of the device-timed loop
    # A full (generation-2) collection of the interpreter's heap takes ~0.75 s once torch is imported and fell
    # into the third timed call on every rank (measured: per-call 914 / 926 / 1675 ms): collect now, and keep
    # the collector out of the timed calls (reference counting still frees every result at once)
    gc.collect()
    gc.disable()
    barrier()
    t0 = time.perf_counter()
    e2e_wall = []
    for _ in range(e2e_steps):
        t_c = time.perf_counter()
        sd = e2e_step()
        e2e_wall.append((time.perf_counter() - t_c) * 1e3)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    gc.enable()
    print(f"[bench] rank {rank}: per-call e2e wall ms {[round(x, 1) for x in e2e_wall]}", file=sys.stderr)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h2d = vpin.nbytes + tpin.nbytes
    d2h = sd["Eigenvalues"].nbytes + sd["Eigenvectors"].nbytes
    parity_e2e = parity_record(sd["Eigenvalues"], golden_spectrum(gkey, args.k), gkey)
    del sd, evec, evec_buf, a_dev, dmesh

    peak, peak_kind = measured_peak()
    configs = {}
    if not args.no_extras:
        if world == 1 and rank == 0:
            configs.update(extras_single_gpu(ctx, peak, mesh))
        try:
            configs["batch_L7"] = batch_record(world)
        except Exception as e:
            configs["batch_L7"] = {"error": repr(e)}
        if world in ROWPART_WORLDS:
            try:
                configs["rowpart"] = rowpart_record(local_rank, world)
            except Exception as e:
                configs["rowpart"] = {"error": repr(e)}
        elif world > 1:
            configs["rowpart"] = {"skipped": f"the row-partitioned solve has been run on {sorted(ROWPART_WORLDS)} GPUs only; a collective "
                                             "that has never executed at this size is not put inside the driver's scaling run"}

    if rank == 0:
        sp = spmm_prof
        class_rate = sp["work"] / (sp["ms"] * 1e-3) / 1e9 if sp["ms"] else 0.0
        total_prof_ms = sum(v["ms"] for v in prof.values())
        classes = {
            k: {"launches": v["launches"], "ms_per_step": v["ms"],
                "rate": v["work"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else None,
                "rate_unit": "GB/s" if k in ("spmm", "col_dots", "elementwise") else "GFLOP/s"}
            for k, v in prof.items()
        }  # fmt: skip
        # roofline of the dominant kernel: algorithmic bytes of its launches in the timed region
        # (12 nnz + 4 (V+1) + 8 m (rows of X + rows of Y [+ rows of B])) / their summed CUDA-event time
        if dom:
            m_dom, dom_f32 = dom["cols"], dom["f32"]
            achieved = dom["work"] / (dom["ms"] * 1e-3) / 1e9
            per_launch_bytes = dom["work"] / dom["launches"]
            kernel = (f"spmm_strip_kernel<{'float' if dom_f32 else 'double'}>, level-0 operator (nnz {nnz:,}) x (n,{m_dom}) block"
                      f"{' of the single-precision multigrid cycle' if dom_f32 else ''}: the SpMM instantiation with the largest "
                      f"summed device time in the timed region ({dom['launches']} launches, {dom['ms'] / args.steps:.1f} ms/step)")
        else:
            m_dom, dom_f32, achieved, per_launch_bytes, kernel = 64, False, 0.0, 0.0, "spmm_strip_kernel (no launches recorded)"
        es_dom = 4 if dom_f32 else 8
        spmm64_bytes = 12 * nnz + 4 * (nv + 1) + 16 * nv * 64
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_dominant_kernel_r2.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if int(tj.get("columns", -1)) == int(m_dom) and tj.get("dtype", "f64") == ("f32" if dom_f32 else "f64"):
                    traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
            except Exception:
                pass
        asm = float(np.median(asm_ms)) if asm_ms else None
        asm_bytes = 4 * mesh.t.shape[1] * nt + 24 * nv + 2 * (12 * nnz + 4 * (nv + 1))
        spmv_bytes = 12 * nnz + 4 * (nv + 1) + 16 * nv
        line = {
            "metric": METRIC, "value": world / (ms_step * 1e-3), "unit": "meshes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "dtype_note": "assembly, operators, LOBPCG blocks, Gram / Rayleigh-Ritz, residuals and results in f64; only the multigrid cycle applied to the residual block (the preconditioner) runs in f32",
            "config": {"workload": desc, "step": "FEM assembly (A, B full) + block-LOBPCG/AMG eigensolve, mesh resident in HBM",
                       "k": args.k, "sigma": -0.01, "tol": "1e-9 scaled residual", "parallelism": f"mesh-parallel x{world}",
                       "l2": "working set per step (S/AS/BS blocks 24 GB at level 9) exceeds the 126 MB L2"},
            "s_per_mesh": ms_step * 1e-3,
            "parity": parity, "parity_e2e": parity_e2e,
            "assembly": {"ms": asm, "gelem_per_s": nt / (asm * 1e-3) / 1e9 if asm else None, "algorithmic_bytes": int(asm_bytes),
                         "roofline_frac": asm_bytes / (asm * 1e-3) / 1e9 / peak if asm else None},
            "spmv": {"ms": spmv_ms, "gb_per_s": spmv_bytes / (spmv_ms * 1e-3) / 1e9, "roofline_frac": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / peak,
                     "renumbered_ms": spmv_ren_ms, "renumbered_roofline_frac": spmv_bytes / (spmv_ren_ms * 1e-3) / 1e9 / peak,
                     "spmm64_renumbered_ms": spmm64_ren_ms, "spmm64_renumbered_roofline_frac": spmm64_bytes / (spmm64_ren_ms * 1e-3) / 1e9 / peak,
                     "note": "ms / roofline_frac: caller's vertex order (what lb_spmm users get); renumbered_*: Morton-cell order used inside the solvers; x, y resident in HBM"},
            "eigs": {"iterations": info["iterations"], "amg_levels": info["amg_levels"], "residual": info["residual"],
                     "amg_setup_ms": info["setup_ms"], "lobpcg_ms": info["solve_ms"]},
            "gpu_launches": int(launches),
            "e2e": {"value": world / e2e_s, "unit": "meshes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind, "kernel": kernel,
                         "algorithmic_bytes": int(per_launch_bytes), "ms_per_launch": dom["ms"] / dom["launches"] if dom else None,
                         "isolated_ms_per_launch": dom_iso_ms,
                         "isolated_frac": ((4 + es_dom) * nnz + 4 * (nv + 1) + 2 * es_dom * nv * m_dom) / (dom_iso_ms * 1e-3) / 1e9 / peak if dom_iso_ms else None,
                         "isolated_note": "y = K x alone (mode 0); the launches in the timed region also read b / write the Chebyshev direction (their bytes are counted)",
                         "dtype": "f32" if dom_f32 else "f64",
                         "spmm_shapes_in_timed_region": [
                             {"columns": s["cols"], "dtype": "f32" if s["f32"] else "f64", "nnz": s["shape"][1], "launches": s["launches"], "ms_per_step": s["ms"] / args.steps,
                              "gb_per_s": s["work"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] else None} for s in spmm_shapes[:12]],
                         "class_in_timed_region": {"launches": sp["launches"], "ms_per_step": sp["ms"] / args.steps, "avg_gb_per_s": class_rate,
                                                   "share_of_step": sp["ms"] / args.steps / ms_step}},
            "kernel_classes": classes,
            "kernel_classes_note": f"CUDA-event totals per kernel class over ONE extra step outside the timed region ({prof_step_ms:.0f} ms with "
                                   f"event pairs around every hot launch; profiled classes sum to {total_prof_ms:.0f} ms)",
            "dense_shapes_in_timed_region": {
                cls: [{"p": r["shape"][0], "q": r["shape"][1], "launches": r["launches"], "ms_per_step": r["ms"],
                       "tflops": r["work"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] else None} for r in rows]
                for cls, rows in dense_shapes.items()},
            "configs": configs,
            "clocks": clk.summary(),
        }  # fmt: skip
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_sample(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not (parity.get("ok", True) and parity_e2e.get("ok", True)):
        print(f"[bench] PARITY FAILURE: {parity} / {parity_e2e}", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
