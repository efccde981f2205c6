import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from nsvf_b200 import synthetic, geometry, clib
from nsvf_b200.clib import _ext
dev = torch.device("cuda:0")
for name in ("C3", "C4"):
    scene = synthetic.make_scene(name)
    pts = torch.from_numpy(scene.points).to(dev)
    centers, children = geometry.build_easy_octree(pts, scene.voxel_size / 2.0)
    centers, children = centers.contiguous(), children.contiguous()
    rs, rd = synthetic.camera_rays(800, 800, 1, radius=4.5, seed=7, device=dev)
    rs = rs.expand_as(rd).reshape(1, -1, 3).contiguous(); rd = rd.reshape(1, -1, 3).contiguous()
    P = scene.max_hits
    t = bench._time(lambda: _ext.svo_intersect(rs, rd, centers, children, scene.voxel_size, P, shared_tree=True), n=5, warm=2)
    idx, dmin, dmax = _ext.svo_intersect(rs, rd, centers, children, scene.voxel_size, P, shared_tree=True)
    t2 = bench._time(lambda: _ext.sort_hits_by_depth(idx.clone(), dmin.clone(), dmax.clone(), 1e4), n=5, warm=2)
    t3 = bench._time(lambda: (idx.clone(), dmin.clone(), dmax.clone()), n=5, warm=2)
    print("%s nodes %d: svo %.3f ms, sort %.3f ms (incl. 3 clones %.3f ms); hits/ray mean %.1f max %d" % (name, centers.shape[0], t, t2, t3, float((idx >= 0).sum(-1).float().mean()), int((idx >= 0).sum(-1).max())))
