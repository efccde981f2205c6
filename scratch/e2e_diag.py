import sys, os, time; sys.path.insert(0, '.')
from nsvf_b200 import blas
blas.use_system_cublas()
import torch, torch.distributed as dist
import bench
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
pipe, scene = bench.build_model(dev, graphs=True)
pipe.field.capture(dev)
model = pipe
if world > 1:
    from torch.nn.parallel import DistributedDataParallel as DDP
    model = DDP(pipe, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True)
opt = torch.optim.Adam([p for p in pipe.parameters() if p.requires_grad], lr=1e-3)
host = bench.make_batches(2, rank, pinned=True)
resident = [tuple(t.to(dev) for t in b) for b in host]
staging = tuple(torch.empty_like(t, device=dev) for t in host[0])
def step(rs, rd, target):
    out = model(rs, rd); loss = bench.loss_fn(out, target)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
def run(n, h2d, item):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); th = 0.0
    for i in range(n):
        if h2d:
            t1 = time.perf_counter()
            for d, s in zip(staging, host[i % 2]): d.copy_(s, non_blocking=True)
            th += time.perf_counter() - t1
            loss = step(*staging)
        else:
            loss = step(*resident[i % 2])
        if item: loss.item()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, th / n * 1e3
run(3, False, False)
for h2d, item in ((False, False), (False, True), (True, False), (True, True)):
    ms, th = run(6, h2d, item)
    print("rank %d world %d  h2d=%s item=%s: %.2f ms/step (host time inside copy_ calls %.2f ms)" % (rank, world, h2d, item, ms, th), flush=True)
# raw copy bandwidth
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for d, s in zip(staging, host[0]): d.copy_(s, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("rank %d raw H2D of %.1f MB: %.2f ms (%.1f GB/s), pinned=%s" % (rank, sum(t.numel() * 4 for t in host[0]) / 1e6, dt * 1e3, sum(t.numel() * 4 for t in host[0]) / dt / 1e9, [t.is_pinned() for t in host[0]]), flush=True)
if world > 1: dist.destroy_process_group()
