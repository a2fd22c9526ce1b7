"""Multi-GPU: independent problems are split contiguously across ranks (one process per GPU).
There is no collective inside an iteration; results are gathered once at the end
(SURVEY.md 8e).  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(B, world, rank):
    """[lo, hi) of the problems rank owns: contiguous, sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, world, rank):
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def gather_to_rank0(local, total_B, group=None):
    """Concatenates per-rank result tensors [B_r, ...] on rank 0 (returns None elsewhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total_B, world, r) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, *local.shape[1:], dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)
