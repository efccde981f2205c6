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
    t4 = bench._time(lambda: _ext.svo_intersect_sorted(rs, rd, centers, children, scene.voxel_size, P, 1e4, shared_tree=True), n=5, warm=2)
    a = _ext.svo_intersect_sorted(rs, rd, centers, children, scene.voxel_size, P, 1e4, shared_tree=True)
    i2, d2, e2 = idx.clone(), dmin.clone(), dmax.clone()
    h2 = _ext.sort_hits_by_depth(i2, d2, e2, 1e4)
    same = torch.equal(a[0], i2) and torch.equal(a[1], d2) and torch.equal(a[2], e2) and torch.equal(a[3], h2)
    index = _ext.SvoIndex(centers, children, scene.voxel_size, shared_tree=True)
    t5 = bench._time(lambda: _ext.svo_intersect_sorted(rs, rd, centers, children, scene.voxel_size, P, 1e4, index=index), n=5, warm=2)
    print("%s svo_intersect_sorted on a prepared octree %.3f ms" % (name, t5))
    print("%s svo_intersect_sorted %.3f ms (identical to traversal + sort: %s)" % (name, t4, same))
    print("%s nodes %d: svo %.3f ms, sort %.3f ms (incl. 3 clones %.3f ms); hits/ray mean %.1f max %d" % (name, centers.shape[0], t, t2, t3, float((idx >= 0).sum(-1).float().mean()), int((idx >= 0).sum(-1).max())))
