import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
dev = torch.device("cuda:0")
pipe, scene = bench.build_model(dev, field="trivial")
params = [p for p in pipe.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)
batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(2, 0, pinned=False)]
def nmalloc():
    s = torch.cuda.memory_stats(dev)
    return s.get("num_device_alloc", 0)
def step(i):
    rs, rd, target = batches[i % 2]
    out = pipe(rs, rd)
    loss = bench.loss_fn(out, target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return out
for i in range(3):
    step(i)
torch.cuda.synchronize()
ts = []
for i in range(30):
    ta = time.perf_counter(); m0 = nmalloc()
    step(i)
    ts.append("%.1f(+%d)" % ((time.perf_counter() - ta) * 1e3, nmalloc() - m0))
torch.cuda.synchronize()
print("unsynced after 3 warmups:", " ".join(ts))
