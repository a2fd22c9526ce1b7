"""Multi-GPU: independent problems are split contiguously across ranks (one process per GPU).
There is no collective inside an iteration; results are gathered once at the end
(SURVEY.md 8e).  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_and_rank(group=None):
    """(world size, rank) of the job, (1, 0) when torch.distributed is not in use."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_bounds(B, world, rank):
    """[lo, hi) of the problems rank owns: contiguous, sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, world, rank):
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def _padded(local, pad):
    buf = torch.zeros(pad, *local.shape[1:], dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    return buf


def gather_to_rank0(local, total_B, group=None):
    """Concatenates per-rank result tensors [B_r, ...] on rank 0 (returns None elsewhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total_B, world, r) for r in range(world)]
    buf = _padded(local, max(hi - lo for lo, hi in sizes))
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def all_gather_problems(tensors, total_B, group=None):
    """The final exchange of a sharded fit: every rank contributes its shard of each result tensor
    ([B_r, ...], e.g. Z, U, K, state) and receives the full [total_B, ...] ones.  All tensors travel in ONE
    collective: they are packed per problem into a byte record, all-gathered, and unpacked."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total_B, world, r) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    Br = tensors[0].shape[0]
    flat = [t.contiguous().reshape(Br, -1).view(torch.uint8) for t in tensors]      # [B_r, bytes_i]
    record = torch.cat(flat, 1)                                                     # one record per problem
    buf = _padded(record, pad)
    out = torch.empty(world * pad, record.shape[1], dtype=torch.uint8, device=record.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    rows = torch.cat([out[r * pad:r * pad + hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
    res, o = [], 0
    for t, f in zip(tensors, flat):
        w = f.shape[1]
        res.append(rows[:, o:o + w].contiguous().view(t.dtype).reshape(total_B, *t.shape[1:]))
        o += w
    return res
