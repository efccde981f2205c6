"""Per-step wall time + cudaMalloc count of the first 40 steps of the C2 hot-path step (stall hunt)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
dev = torch.device("cuda:0")
field = sys.argv[1] if len(sys.argv) > 1 else "trivial"
pipe, scene = bench.build_model(dev, field=field)
params = [p for p in pipe.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)
batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(2, 0, pinned=False)]
def step(i):
    rs, rd, target = batches[i % 2]
    out = pipe(rs, rd)
    loss = bench.loss_fn(out, target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return out
def nmalloc():
    s = torch.cuda.memory_stats(dev)
    return s.get("num_device_alloc", 0), s.get("num_device_free", 0), s.get("num_alloc_retries", 0), s.get("reserved_bytes.all.current", 0) >> 20
line = []
for i in range(40):
    a = nmalloc(); torch.cuda.synchronize(); t0 = time.perf_counter()
    step(i)
    torch.cuda.synchronize(); t1 = time.perf_counter(); b = nmalloc()
    line.append("%d:%.1fms(+%dmalloc,%dfree,res%dMB)" % (i, (t1 - t0) * 1e3, b[0] - a[0], b[1] - a[1], b[3]))
print(" ".join(line))
