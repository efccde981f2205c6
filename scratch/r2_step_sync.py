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
def step(i):
    rs, rd, target = batches[i % 2]
    out = pipe(rs, rd)
    loss = bench.loss_fn(out, target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return out
for i in range(5):
    step(i)
def run(n, sync, label):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ts = []
    for i in range(n):
        ta = time.perf_counter()
        step(i)
        if sync:
            torch.cuda.synchronize()
        ts.append((time.perf_counter() - ta) * 1e3)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("%s: %.2f ms/step device, %.2f wall; host per step: %s" % (label, e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n, " ".join("%.1f" % t for t in ts)))
run(12, True, "synced")
run(12, False, "unsynced")
run(12, True, "synced")
os.environ["NSVF_RENDER_GENERAL"] = "1"
pipe.padded_samples = True
run(12, False, "general unsynced")
run(12, True, "general synced")
