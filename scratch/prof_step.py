import sys, os; sys.path.insert(0, '.')
from nsvf_b200 import blas
blas.use_system_cublas()
import torch, time, cProfile, pstats, io
import bench
dev = torch.device('cuda:0')
pipe, scene = bench.build_model(dev)
opt = torch.optim.Adam([p for p in pipe.parameters() if p.requires_grad], lr=1e-3)
host = bench.make_batches(2, 0, pinned=False)
res = [tuple(t.to(dev) for t in b) for b in host]
def step(i):
    rs, rd, target = res[i % 2]
    out = pipe(rs, rd)
    loss = bench.loss_fn(out, target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss
for i in range(3): step(i)
torch.cuda.synchronize()
# enqueue time vs device time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(5): step(i)
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("per step: host enqueue %.2f ms, host+drain %.2f ms, device events %.2f ms" % (t_enq / 5 * 1e3, t_all / 5 * 1e3, e0.elapsed_time(e1) / 5))
# forward / backward / opt split (with syncs)
def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
rs, rd, target = res[0]
out, t_f = timed(lambda: pipe(rs, rd))
loss, t_l = timed(lambda: bench.loss_fn(out, target))
opt.zero_grad(set_to_none=True)
_, t_b = timed(lambda: loss.backward())
_, t_o = timed(lambda: opt.step())
print("forward %.2f ms, loss %.2f, backward %.2f, adam %.2f" % (t_f, t_l, t_b, t_o))
pr = cProfile.Profile(); pr.enable()
for i in range(5): step(i)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): step(i)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(e.self_device_time_total for e in ka) / 3 / 1e3
print("sum of device kernel time per step: %.2f ms" % tot)
rows = sorted(ka, key=lambda e: -e.self_device_time_total)[:28]
for e in rows:
    print("%8.3f ms %5d  %s" % (e.self_device_time_total / 3 / 1e3, e.count // 3, e.key[:110]))
