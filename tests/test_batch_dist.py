"""Host-side logic of the mesh-parallel batch (lapy_b200/batch.py) on CPU: sharding and the
world_size-2 gather over gloo (the device compute is replaced by a deterministic stand-in)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lapy_b200.batch import batched_shapedna, row_partition, shard_indices


def test_shard_indices_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 512):
            shards = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _fake_compute(mesh, k, lump):
    return np.arange(k, dtype=np.float64) * (mesh + 1) + (0.5 if lump else 0.0)


def test_serial_batch_matches_loop():
    out = batched_shapedna(lambda i: i, n_meshes=5, k=4, compute=_fake_compute)
    assert out.shape == (5, 4)
    for i in range(5):
        np.testing.assert_array_equal(out[i], _fake_compute(i, 4, False))
    local = batched_shapedna([3, 4], k=3, lump=True, compute=_fake_compute, gather=False)
    assert set(local) == {0, 1}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []

        def compute(mesh, k, lump):
            seen.append(mesh)
            return _fake_compute(mesh, k, lump)

        out = batched_shapedna(lambda i: i, n_meshes=7, k=5, compute=compute)
        q.put((rank, out, sorted(seen)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([_fake_compute(i, 5, False) for i in range(7)])
    for rank, out, seen in res:
        np.testing.assert_array_equal(out, ref)  # every rank holds the full table
        assert seen == list(range(rank, 7, 2))  # and computed only its own shard


def test_row_partition_covers_all_rows():
    for n in (1, 7, 40962, 2621442, 1771561):
        for world in (1, 2, 4, 8):
            parts = row_partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) == -(-n // world) and all(s >= 0 for s in sizes)


def test_worker_threads_process_every_mesh_once():
    import threading
    import time

    seen, names = [], set()
    lock = threading.Lock()

    def compute(mesh, k, lump):
        time.sleep(0.01)
        with lock:
            seen.append(mesh)
            names.add(threading.current_thread().name)
        return _fake_compute(mesh, k, lump)

    out = batched_shapedna(lambda i: i, n_meshes=23, k=4, compute=compute, workers=4)
    assert sorted(seen) == list(range(23))
    assert len(names) == 4 and all(n.startswith("lapy-b200-batch-") for n in names)
    for i in range(23):
        np.testing.assert_array_equal(out[i], _fake_compute(i, 4, False))

    got_ctx = []

    def compute4(mesh, k, lump, ctx):  # a custom compute may take the worker's context ...
        got_ctx.append(ctx)
        return _fake_compute(mesh, k, lump)

    batched_shapedna([0, 1, 2], k=3, compute=compute4)  # ... which is None on the serial path
    assert got_ctx == [None, None, None]

    def boom(mesh, k, lump):
        if mesh == 5:
            raise RuntimeError("mesh 5 is broken")
        return _fake_compute(mesh, k, lump)

    with pytest.raises(RuntimeError, match="mesh 5"):
        batched_shapedna(lambda i: i, n_meshes=12, k=4, compute=boom, workers=3)
    with pytest.raises(ValueError, match="workers"):
        batched_shapedna([0], k=2, compute=_fake_compute, workers=0)
