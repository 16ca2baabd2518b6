"""Multi-process host logic on CPU (gloo, world_size 2): the frame partition is disjoint and complete, and the pose gather
returns the frames in global order on every rank — the N>1 path of bench.py minus the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpg_monocular_pose_estimator_b200.sharding import shard_range, gather_poses


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 8193):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                got.extend(range(a, b))
            assert got == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_frames, rank, world)
    # stand-in for the per-rank device result: pose f carries its global frame index
    local = torch.arange(a, b, dtype=torch.float64).reshape(-1, 1).repeat(1, 16)
    allp = gather_poses(local, world, dist)
    q.put((rank, allp[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_pose_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames = 11                      # ragged: 6 + 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for r in range(2):
        assert res[r] == [float(i) for i in range(n_frames)]
