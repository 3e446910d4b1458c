"""BASELINE.json config 5 (BrainPrint-style batch): N perturbed level-7 icospheres (163,842 vertices),
ShapeDNA k=50 each, mesh-parallel over the ranks of a torchrun job and `--workers` concurrent
contexts per GPU.  Prints meshes/s.

    python tools/bench_batch.py --meshes 32 --workers 1
    python tools/bench_batch.py --meshes 32 --workers 4
    torchrun --nproc-per-node 8 tools/bench_batch.py --meshes 512 --workers 4
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapy_b200 import mesh as M  # noqa: E402
from lapy_b200.batch import batched_shapedna  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--meshes", type=int, default=32)
ap.add_argument("--workers", type=int, default=1)
ap.add_argument("--level", type=int, default=7)
ap.add_argument("--k", type=int, default=50)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

cache = {}


def factory(i):
    if i not in cache:  # built before the timed region (host-side mesh generation is not the path)
        cache[i] = M.perturbed_sphere(args.level, seed=i)
    return cache[i]


mine = range(rank, args.meshes, world)
for i in mine:
    factory(i)
batched_shapedna(factory, n_meshes=min(args.meshes, max(2, args.workers) * world), k=args.k, workers=args.workers)  # warm-up: every worker context once
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ev = batched_shapedna(factory, n_meshes=args.meshes, k=args.k, workers=args.workers)
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"metric": "batch_shapedna_meshes_per_s", "value": args.meshes / dt, "seconds": dt, "meshes": args.meshes,
                      "vertices_per_mesh": int(factory(rank).v.shape[0]), "k": args.k, "n_gpus": world, "workers_per_gpu": args.workers,
                      "ev1_range": [float(ev[:, 1].min()), float(ev[:, 1].max())]}))  # fmt: skip
if world > 1:
    dist.destroy_process_group()
