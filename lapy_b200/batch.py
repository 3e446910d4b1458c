"""Mesh-parallel batches (BASELINE.json config 5: BrainPrint-style ShapeDNA over many surfaces).

The path shards by mesh: every unit (one surface) is independent, so rank r of a
``torch.distributed`` job (one process per GPU, torchrun) processes meshes ``r, r+W, r+2W, ...`` on
its own device and context; there is no data-path collective - only the k eigenvalues per mesh
are gathered at the end.  The reference has no equivalent (single process).
"""

from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment mesh i -> rank i mod world."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return list(range(rank, n_items, world))


# worker contexts (one CUDA stream + cuSOLVER handle + work space each) are kept between calls: creating one
# costs ~0.1 s (handles, pinned staging), comparable to a whole level-7 solve
_ctx_pool: dict = {}
_ctx_lock = __import__("threading").Lock()


def _take_context(device: int):
    from . import _lib

    with _ctx_lock:
        free = _ctx_pool.setdefault(device, [])
        if free:
            return free.pop()
    return _lib.Context(device)


def _return_context(ctx):
    with _ctx_lock:
        _ctx_pool.setdefault(ctx.device, []).append(ctx)


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def batched_shapedna(
    meshes: Sequence | Callable[[int], object],
    n_meshes: int | None = None,
    k: int = 50,
    lump: bool = False,
    compute: Callable | None = None,
    gather: bool = True,
    workers: int = 1,
):
    """ShapeDNA eigenvalues of a batch of meshes, sharded over the ranks of the current process
    group (or run serially without one).

    ``meshes``: a sequence of geometries or a factory ``i -> geometry`` (with ``n_meshes``), so
    that a rank only materialises its own shard.  ``compute(mesh, k, lump) -> (k,) eigenvalues``
    defaults to :func:`lapy_b200.shapedna.compute_shapedna` on this rank's GPU.
    Returns an ``(n_meshes, k)`` array on every rank (``gather=True``) or ``{index: (k,)}`` of
    the local shard.

    ``workers`` > 1 processes that many meshes of the shard concurrently on this rank's GPU, one
    host thread and one library context (= one CUDA stream, one cuSOLVER handle) each: at ~150k
    vertices a mesh is launch- and latency-bound (small kernels, the dense Rayleigh-Ritz solve, host
    synchronisations), so several streams fill the device.  The default ``compute`` then receives its
    worker's context; a custom ``compute`` may take it as a fourth argument.
    """
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    if callable(meshes):
        if n_meshes is None:
            raise ValueError("n_meshes is required with a mesh factory")
        get = meshes
    else:
        n_meshes = len(meshes)
        get = meshes.__getitem__
    if workers < 1:
        raise ValueError("workers must be >= 1")
    takes_ctx = False
    if compute is None:
        from .solver import Solver

        def compute(mesh, k, lump, ctx=None):
            return Solver(mesh, lump=lump, ctx=ctx).eigs(k=k, vectors=False)[0]  # the batch returns eigenvalues only

        takes_ctx = True
    else:
        import inspect

        try:
            takes_ctx = len(inspect.signature(compute).parameters) >= 4
        except (TypeError, ValueError):
            takes_ctx = False

    def one(i, ctx):
        ev = compute(get(i), k, lump, ctx) if takes_ctx else compute(get(i), k, lump)
        ev = np.asarray(ev, dtype=np.float64)
        if ev.shape != (k,):
            raise ValueError(f"compute returned shape {ev.shape}, expected ({k},)")
        return ev

    mine = shard_indices(n_meshes, rank, world)
    local = {}
    if workers == 1 or len(mine) <= 1:
        for i in mine:
            local[i] = one(i, None)
    else:
        import queue
        import threading

        todo: queue.SimpleQueue = queue.SimpleQueue()
        for i in mine:
            todo.put(i)
        errors: list[BaseException] = []
        lock = threading.Lock()

        def run():
            ctx = None
            if takes_ctx:
                try:
                    ctx = _take_context(int(os.environ.get("LAPY_B200_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
                except BaseException as e:  # noqa: BLE001 - no device, library missing: re-raised by the caller
                    with lock:
                        errors.append(e)
                    return
            try:
                work(ctx)
            finally:
                if ctx is not None:
                    _return_context(ctx)

        def work(ctx):
            while not errors:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    ev = one(i, ctx)
                except BaseException as e:  # noqa: BLE001 - re-raised on the calling thread
                    with lock:
                        errors.append(e)
                    return
                with lock:
                    local[i] = ev

        threads = [threading.Thread(target=run, name=f"lapy-b200-batch-{w}") for w in range(min(workers, len(mine)))]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    if not gather:
        return local
    out = np.zeros((n_meshes, k), np.float64)
    if dist is None:
        for i, ev in local.items():
            out[i] = ev
        return out
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dev == "cuda":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    buf = torch.zeros((n_meshes, k), dtype=torch.float64, device=dev)
    for i, ev in local.items():
        buf[i] = torch.from_numpy(ev).to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # disjoint rows: a sum is a gather
    return buf.cpu().numpy()


def row_partition(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row blocks of the row-partitioned single-mesh mode (csrc/eigs.cu lobpcg_dist):
    ``rows_per_rank = ceil(n / world)``, rank r owns ``[r*rpr, min(n, (r+1)*rpr))`` of the
    locality-renumbered operator."""
    if world < 1:
        raise ValueError("world must be >= 1")
    rpr = (n + world - 1) // world
    return [(min(n, r * rpr), min(n, (r + 1) * rpr)) for r in range(world)]
