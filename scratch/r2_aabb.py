import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from nsvf_b200 import synthetic
from nsvf_b200.clib import _ext
dev = torch.device("cuda:0")
for name in ("C3", "C4"):
    scene = synthetic.make_scene(name)
    pts = torch.from_numpy(scene.points).to(dev); pts[:, 0] += scene.voxel_size / 10
    rs, rd = synthetic.camera_rays(800, 800, 1, radius=4.5, seed=7, device=dev)
    rs = rs.expand_as(rd).reshape(1, -1, 3).contiguous(); rd = rd.reshape(1, -1, 3).contiguous()
    t = bench._time(lambda: _ext.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, scene.max_hits, 1e4, shared_points=True), n=5, warm=2)
    t2 = bench._time(lambda: _ext.aabb_hit_mask(rs, rd, pts, scene.voxel_size, shared_points=True), n=5, warm=2)
    print("SMEM_NODES=%s LIST_CAP=%s %s: sorted %.3f ms, any-hit %.3f ms" % (os.environ.get("NSVF_AABB_SMEM_NODES"), os.environ.get("NSVF_AABB_LIST_CAP"), name, t, t2))
