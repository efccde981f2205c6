"""C2 training step with the stand-in field (hot path only): device time, host profile, launch list under ncu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
dev = torch.device("cuda:0")
field = sys.argv[1] if len(sys.argv) > 1 else "trivial"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
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
for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(n):
    out = step(i)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("step %s: %.3f ms device, %.3f ms wall, ae=%d" % (field, e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n, out["ae"]))
if os.environ.get("NSVF_PROFILE_PY"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for i in range(5):
        step(i)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(40)
