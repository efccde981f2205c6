"""C3 frame, hot path only (trivial field): per-frame device time, host time and (under ncu) the launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
dev = torch.device("cuda:0")
from nsvf_b200 import synthetic
from nsvf_b200.field import TrivialField
rs, rd = synthetic.camera_rays(800, 800, 1, radius=4.5, seed=7, device=dev)
rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
field = sys.argv[1] if len(sys.argv) > 1 else "trivial"
pipe, scene = bench.build_model(dev, "C3", train=False, field=field, tolerance=0.01, chunk=512, sigma_bias=2.0)
pipe.lazy_sampling = os.environ.get("NSVF_LAZY", "1") == "1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
with torch.no_grad():
    for _ in range(2):
        out = pipe(rs, rd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = pipe(rs, rd)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
print("frame %s: %.3f ms device, %.3f ms wall, ae=%d hits=%d" % (field, e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n, out["ae"], int(out["hits"].sum())))
if os.environ.get("NSVF_PROFILE_PY"):
    import cProfile, pstats
    with torch.no_grad():
        pr = cProfile.Profile(); pr.enable()
        for _ in range(3):
            out = pipe(rs, rd)
        torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
