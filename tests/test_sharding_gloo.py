"""N>1 host logic on CPU: two gloo ranks shard a batch of problems contiguously, work on their
slice independently (no collective in the 'iteration'), and rank 0 gathers the results."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pddp_b200.sharding import all_gather_problems, gather_to_rank0, shard, shard_bounds, world_and_rank


def _worker(rank, world, port, B, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(B * 3 * 2, dtype=torch.float32).reshape(B, 3, 2)
    mine = shard(full, world, rank)
    lo, hi = shard_bounds(B, world, rank)
    assert mine.shape[0] == hi - lo
    result = mine * 2 + 1                      # stands in for the per-rank solver
    got = gather_to_rank0(result, B)
    if rank == 0:
        torch.save(got, out_path)
    else:
        assert got is None
    # the product's final exchange (iLQRController.fit under torchrun): Z, U, K, state in ONE all-gather
    assert world_and_rank() == (world, rank)
    Z = mine.double() + 0.5
    state = torch.arange(lo, hi, dtype=torch.int32)
    fZ, fU, fs = all_gather_problems([Z, result, state], B)
    assert torch.equal(fZ, full.double() + 0.5) and torch.equal(fU, full * 2 + 1)
    assert torch.equal(fs, torch.arange(B, dtype=torch.int32))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for B in (1, 7, 4096, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather(tmp_path):
    B, world = 7, 2
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, 29731, B, out), nprocs=world, join=True)
    got = torch.load(out)
    want = torch.arange(B * 3 * 2, dtype=torch.float32).reshape(B, 3, 2) * 2 + 1
    assert torch.equal(got, want)
