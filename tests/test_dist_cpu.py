"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding covers every ray exactly once, the
keep-mask all-gather equals the single-rank mask, the gradient all-reduce averages."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from nsvf_b200 import dist as nd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ray sharding: union of shards == all rays, disjoint
        V, P = 5, 37
        rd = torch.arange(V * P * 3, dtype=torch.float32).view(1, V, P, 3)
        rs = torch.zeros(1, V, 1, 3)
        _, mine = nd.shard_rays(rs, rd, rank, world, dim=1)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        assert torch.equal(torch.cat(parts, 1), rd)
        # keep-mask all-gather with an uneven split
        n = 1001
        full = (torch.arange(n) * 7919 % 13 < 6)
        lo, hi = nd.shard_range(n, rank, world)
        got = nd.allgather_keep_mask(full[lo:hi], n, rank, world)
        assert got.dtype == torch.bool and torch.equal(got, full)
        # gradient all-reduce (mean), embedding + dense parameters
        emb = torch.nn.Parameter(torch.zeros(50, 32))
        w = torch.nn.Parameter(torch.zeros(8, 8))
        emb.grad = torch.full_like(emb, float(rank + 1))
        w.grad = torch.full_like(w, float(10 * (rank + 1)))
        nd.allreduce_grads([emb, w], world)
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(emb.grad, torch.full_like(emb, mean)) and torch.allclose(w.grad, torch.full_like(w, 10 * mean))
        frames = nd.gather_frames(["f%d_%d" % (rank, i) for i in range(2)], rank, world)
        if rank == 0:
            assert frames == ["f0_0", "f0_1", "f1_0", "f1_1"]
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_range_partitions():
    from nsvf_b200 import dist as nd
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [nd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
