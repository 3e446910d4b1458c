"""Row-partitioned (multi-GPU, NCCL) eigensolve: 2 ranks cooperate on ONE mesh and must reproduce
the reference spectrum like the single-GPU path.  Needs 2 GPUs; skipped otherwise."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, env):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))  # fmt: skip
    os.environ.update(env)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lapy_b200
        from lapy_b200 import _lib
        from lapy_b200 import mesh as M

        ctx = _lib.Context(rank)
        ctx.init_row_partition()
        mesh = M.icosphere(6)
        fem = lapy_b200.Solver(mesh, ctx=ctx)
        ev, evec = fem.eigs(k=50)
        b = fem.mass
        orth = float(np.abs(evec.T @ (b @ evec) - np.eye(50)).max())
        res = float(np.abs(fem.stiffness @ evec - (b @ evec) * ev).max() / np.abs(b @ evec).max())
        q.put((rank, ev, orth, res, fem.last_info))
        ctx.leave_row_partition()
    finally:
        dist.destroy_process_group()


def test_two_rank_row_partitioned_eigs():
    """Rows of the locality-numbered operator split over 2 ranks: NCCL halo exchange in every SpMM,
    all-reduced Gram matrices, hierarchy of the full operator applied column-parallel.  Validated on
    2 GPUs in round 2 (level 6 here; level 9: 40 iterations, 1.14 s; 121^3 tets: 46 iterations, 1.20 s)."""
    import torch
    import torch.multiprocessing as mp
    from conftest import load_golden

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, {})) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = load_golden("spectra")["ico6_k50"]
    for rank, ev, orth, resid, info in res:
        assert np.all(np.abs(ev[1:] - ref[1:]) <= 1e-8 * ref[1:]) and abs(ev[0]) < 1e-8, (rank, ev[:4], ref[:4])
        assert orth < 1e-9 and resid < 1e-6
    np.testing.assert_array_equal(res[0][1], res[1][1])  # both ranks hold identical results
    assert res[0][4]["iterations"] <= 70
