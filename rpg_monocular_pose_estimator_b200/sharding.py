"""Frame / stream partition used by the multi-GPU driver: frames (cold mode) or streams (tracking)
are independent, so rank r of G takes a contiguous block and no data-path collective exists; only the fixed-size pose
records are gathered.  Pure host logic, usable with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced block of `n_items` for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_poses(local_poses, world: int, dist=None):
    """all-gather of per-rank [n_local, 16] pose tensors of possibly different n_local; returns the global [N,16] tensor
    in frame order on every rank.  `dist` = torch.distributed (process group already initialised)."""
    import torch
    if world == 1 or dist is None:
        return local_poses
    n_local = torch.tensor([local_poses.shape[0]], dtype=torch.int64, device=local_poses.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    n_max = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((n_max, 16), dtype=local_poses.dtype, device=local_poses.device)
    padded[:local_poses.shape[0]] = local_poses
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:int(c.item())] for o, c in zip(out, counts)], dim=0)
