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


def gather_records(local, world: int, dist=None, out=None):
    """all-gather of equal-sized per-rank record blocks (`local`: [n_local, rec_bytes] uint8, or any [n_local, k] tensor) into
    [world * n_local, k] in rank order; `out` may be a preallocated destination.  This is the collective of the throughput bench
    (NCCL gather of the mpe_result records); ragged blocks go through gather_poses' padding scheme."""
    import torch
    if world == 1 or dist is None:
        return local
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def record_checksum(block):
    """Order-sensitive 64-bit checksum of a block of records (wrapping int64 arithmetic on the raw bytes)."""
    import torch
    b = block.contiguous().view(torch.uint8).reshape(-1).to(torch.int64)
    w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) % 65521) + 1
    return (b * w).sum()


def verify_gather(gathered, local, rank: int, world: int, dist=None) -> bool:
    """True iff the gathered table holds every rank's block unchanged: the own block is compared byte for byte, the others
    through a checksum exchange (each rank publishes the checksum of what it sent)."""
    import torch
    n = local.shape[0]
    if world == 1 or dist is None:
        return bool(torch.equal(gathered, local))
    mine = record_checksum(local).reshape(1)
    sums = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sums, mine)
    ok = bool(torch.equal(gathered[rank * n:(rank + 1) * n], local))
    for r in range(world):
        ok = ok and int(record_checksum(gathered[r * n:(r + 1) * n]).item()) == int(sums[r].item())
    return ok
