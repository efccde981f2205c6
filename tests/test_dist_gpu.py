"""Multi-GPU (NCCL) tests of the two exchange steps of the path: the gradient all-reduce and the keep-mask all-gather of
voxel-sharded pruning.  Needs >= 2 GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`);
skipped on a single-GPU box, where tests/test_dist_cpu.py covers the same host logic on gloo."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from nsvf_b200 import dist as nd, synthetic
    from nsvf_b200.encoder import SparseVoxelEncoder
    try:
        # 1) gradient all-reduce (mean) over NCCL == mean of the per-rank gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(1000, 32, device=dev)), torch.nn.Parameter(torch.zeros(257, device=dev))]
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1)) + torch.arange(p.numel(), device=dev).view_as(p) * 1e-3
        expect = [sum(torch.full_like(p, float(r + 1)) for r in range(world)) / world
                  + torch.arange(p.numel(), device=dev).view_as(p) * 1e-3 for p in params]
        nd.allreduce_grads(params, world)
        ok_grad = all(torch.allclose(p.grad, e, rtol=1e-6, atol=1e-6) for p, e in zip(params, expect))
        # 2) voxel-sharded pruning: the all-gathered keep mask equals the mask one rank computes alone, bit for bit
        scene = synthetic.make_scene("C1")
        enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=60).to(dev)
        with torch.no_grad():
            enc.values.weight.copy_(torch.from_numpy(scene.values).to(dev))
        field = lambda inp, outputs: {"sigma": inp["emb"][:, 0] * 8 - 0.2}
        keep0 = enc.keep.clone()
        enc.pruning(field, th=0.5, voxel_shard=(rank, world), bits=8)
        sharded = enc.keep.clone()
        enc.keep.copy_(keep0)
        enc.invalidate_geometry_cache()
        enc.pruning(field, th=0.5, bits=8)
        ok_mask = bool(torch.equal(sharded, enc.keep)) and 0 < int(sharded.sum()) < sharded.numel()
        # every rank ends with the same mask
        ref = sharded.clone()
        dist.broadcast(ref, src=0)
        ok_same = bool(torch.equal(ref, sharded))
        out[rank] = (ok_grad, ok_mask, ok_same)
    finally:
        dist.destroy_process_group()


def test_nccl_grad_allreduce_and_keep_mask_allgather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] == (True, True, True) for r in range(world)), dict(out)
