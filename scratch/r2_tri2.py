import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from nsvf_b200 import _lib
dev = torch.device("cuda:0")
L, p = _lib.load(), _lib.ptr
scene, vox, xyz, feats, pts, values, runs = bench.real_sample_stream(dev)
M = vox.numel()
out = torch.empty(M, 32, device=dev); gv = torch.zeros_like(values)
st = torch.cuda.current_stream().cuda_stream
f = bench._time(lambda: L.nsvf_trilinear_embed_fwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size, p(out)), n=5, warm=2)
b = bench._time(lambda: L.nsvf_trilinear_embed_bwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size, p(out), p(gv), None), n=5, warm=2)
print("VARIANT=%s SNAP=%s BWD=%s PERM=%s BPS=%s: fwd %.3f ms (%.3f)  bwd %.3f ms (%.3f)" % (os.environ.get("NSVF_TRI_VARIANT"), os.environ.get("NSVF_TRI_SNAP"), os.environ.get("NSVF_TRI_BWD"), os.environ.get("NSVF_TRI_PERM"), os.environ.get("NSVF_TRI_BPS"), f, M*144/f/1e6/6545.6, b, M*144/b/1e6/6545.6))
