"""Quick device timings of the fused ops at C3 scale (not a pytest file)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nsvf_b200 import synthetic, ops, _lib
L = _lib.load(); p = _lib.ptr
dev = torch.device("cuda:0")
scene = synthetic.make_scene("C3")
pts = torch.from_numpy(scene.points).to(dev); feats = torch.from_numpy(scene.feats).int().to(dev)
values = torch.from_numpy(scene.values).to(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
g = torch.Generator(device=dev).manual_seed(0)
RUN = int(sys.argv[2]) if len(sys.argv) > 2 else 6      # consecutive samples per voxel (step = voxel/8 -> 6..10)
runs = torch.randint(0, scene.n, (M // RUN + 1,), device=dev, generator=g)
vox = runs.repeat_interleave(RUN)[:M].int().contiguous()
xyz = (pts[vox.long()] + (torch.rand(M, 3, device=dev, generator=g) - 0.5) * scene.voxel_size).contiguous()
out = torch.empty(M, 32, device=dev); gout = torch.randn(M, 32, device=dev); gv = torch.zeros_like(values)
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
f = lambda: L.nsvf_trilinear_embed_fwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size, p(out))
b = lambda: L.nsvf_trilinear_embed_bwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size, p(gout), p(gv), None)
print("run length", RUN, "NSVF_TRI_SNAP", os.environ.get("NSVF_TRI_SNAP"), "NSVF_TRI_BWD", os.environ.get("NSVF_TRI_BWD"))
ms = t(f); print("trilinear fwd  M=%d: %.3f ms  %.0f GB/s (144 B/sample)" % (M, ms, M * 144 / ms / 1e6))
ms = t(b); print("trilinear bwd  M=%d: %.3f ms  %.0f GB/s (144 B/sample)" % (M, ms, M * 144 / ms / 1e6))
if os.environ.get("TRI_ONLY"): sys.exit(0)
B, K = 262144, 256
fe = torch.rand(B, K, device=dev) * 0.1; tex = torch.rand(B, K, 3, device=dev); dep = torch.rand(B, K, device=dev)
probs = torch.empty(B, K, device=dev); od = torch.empty(B, device=dev); om = torch.empty(B, device=dev); oc = torch.empty(B, 3, device=dev)
cf = lambda: L.nsvf_composite_fwd(st, B, K, p(fe), p(tex), p(dep), p(probs), p(od), p(om), p(oc))
ms = t(cf); print("composite fwd [%d,%d]: %.3f ms  %.0f GB/s (28 B/sample)" % (B, K, ms, (B * K * 28 + B * 20) / ms / 1e6))
gfe = torch.empty(B, K, device=dev); gtex = torch.empty(B, K, 3, device=dev)
gd = torch.randn(B, device=dev); gm = torch.randn(B, device=dev); gc = torch.randn(B, 3, device=dev)
cb = lambda: L.nsvf_composite_bwd(st, B, K, p(fe), p(tex), p(dep), None, p(gd), p(gm), p(gc), p(gfe), p(gtex))
ms = t(cb); print("composite bwd [%d,%d]: %.3f ms  %.0f GB/s (36 B/sample)" % (B, K, ms, (B * K * 36 + B * 20) / ms / 1e6))
