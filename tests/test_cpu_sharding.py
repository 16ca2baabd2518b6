"""Multi-process host logic on CPU (gloo, world_size 2): the frame partition is disjoint and complete, and the pose gather
returns the frames in global order on every rank — the N>1 path of bench.py minus the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpg_monocular_pose_estimator_b200.sharding import shard_range, gather_poses, gather_records, verify_gather


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
    # the record gather of bench.py: equal blocks of raw result records, verified by checksum exchange
    rec = (torch.arange(4 * 968, dtype=torch.int64) * (rank + 3) % 251).to(torch.uint8).reshape(4, 968)
    table = gather_records(rec, world, dist)
    ok = verify_gather(table, rec, rank, world, dist)
    bad = table.clone(); bad[(1 - rank) * 4, 5] ^= 1                 # one flipped bit in the OTHER rank's block must be noticed
    ok_bad = verify_gather(bad, rec, rank, world, dist)
    q.put((rank, allp[:, 0].tolist(), table.shape[0], ok, ok_bad))
    dist.barrier()
    dist.destroy_process_group()


def test_pose_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames = 11                      # ragged: 6 + 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs: p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    res = {g[0]: g[1] for g in got}
    assert all(g[2] == 8 and g[3] is True and g[4] is False for g in got), got
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for r in range(2):
        assert res[r] == [float(i) for i in range(n_frames)]
