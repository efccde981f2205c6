"""Live per-kernel times inside the C3 hot-path frame (bench.frame_rooflines)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from nsvf_b200 import synthetic
dev = torch.device("cuda:0")
rs, rd = synthetic.camera_rays(800, 800, 1, radius=4.5, seed=7, device=dev)
rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
pipe, scene = bench.build_model(dev, "C3", train=False, field="trivial", tolerance=0.01, chunk=512, sigma_bias=2.0)
with torch.no_grad():
    for _ in range(2): pipe(rs, rd)
t = bench._time(lambda: pipe(rs, rd), n=5, warm=1)
print("frame %.3f ms" % t)
for r in bench.frame_rooflines(dev, pipe, rs, rd, 6545.6, "measured", t):
    print("%-32s %3d launches %7.3f ms/frame  %8.1f GB/s  frac %.3f" % (r["kernel"], r["launches_per_frame"], r["ms_per_frame"], r["achieved"], r["frac"]))
