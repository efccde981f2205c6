"""Stall hunt at N=2: the e2e loop of bench.py (H2D from pinned memory + step + loss.item()) with a per-step breakdown."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "base"
pipe, scene = bench.build_model(dev, graphs=True)
pipe.field.capture(dev)
params = [p for p in pipe.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)
flat = torch.zeros(sum(p.numel() for p in params), device=dev); off = 0
for p in params:
    p.grad = flat[off: off + p.numel()].view_as(p); off += p.numel()
host = bench.make_batches(2, rank, pinned=True)
staging = tuple(torch.empty_like(t, device=dev) for t in host[0])
resident = [tuple(t.to(dev) for t in b) for b in host]
def step(rs, rd, target):
    out = pipe(rs, rd)
    loss = bench.loss_fn(out, target)
    flat.zero_()
    loss.backward()
    if world > 1 and mode != "nonccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    opt.step()
    return loss
for i in range(8):
    step(*resident[i % 2])
torch.cuda.synchronize()
if world > 1: dist.barrier()
slow = []
import faulthandler
tb_file = open(os.path.join(ROOT, "gpurun_out", "r2_stall_tb_rank%d.txt" % rank), "a")
def cgstat():
    for f in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.stat"):
        try:
            return {l.split()[0]: int(l.split()[1]) for l in open(f)}
        except Exception:
            pass
    return {}
cg0 = cgstat()
if rank == 0:
    try:
        print("cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip(), "nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
    except Exception as e:
        print("cpu.max unavailable", e, "nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
evs = []
t_all0 = time.perf_counter()
for i in range(80):
    faulthandler.dump_traceback_later(0.08, exit=False, file=tb_file)
    m0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    t0 = time.perf_counter(); c0 = time.process_time()
    ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ea.record()
    if mode == "nocopy":
        args = resident[i % 2]
    else:
        for d, s in zip(staging, host[i % 2]):
            d.copy_(s, non_blocking=True)
        args = staging
    eb.record()
    t1 = time.perf_counter()
    loss = step(*args)
    ec.record()
    t2 = time.perf_counter()
    if mode == "noitem":
        torch.cuda.current_stream().synchronize()
    else:
        loss.item()
    t3 = time.perf_counter()
    faulthandler.cancel_dump_traceback_later()
    if (t3 - t0) > 0.06:
        slow.append("mallocs +%d" % (torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - m0))
        torch.cuda.synchronize()
        slow.append("step %d: total %.0f ms = copies %.1f + enqueue %.1f + wait %.1f | cpu time of main thread+children %.0f ms | GPU: copies %.1f ms, step %.1f ms" % (
            i, (t3 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.process_time() - c0) * 1e3, ea.elapsed_time(eb), eb.elapsed_time(ec)))
tot = (time.perf_counter() - t_all0) * 1e3 / 80
cg1 = cgstat()
print("rank %d mode %s: %.1f ms/step; cgroup throttled periods +%s, throttled usec +%s; slow steps: %s" % (
    rank, mode, tot, cg1.get("nr_throttled", 0) - cg0.get("nr_throttled", 0), cg1.get("throttled_usec", 0) - cg0.get("throttled_usec", 0),
    "; ".join(slow) or "none"), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
