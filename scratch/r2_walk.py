"""Lattice walk vs hierarchy kernels: sorted / index-order / any-hit intersection on the C2, C3, C4 scenes (device time)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from nsvf_b200 import synthetic
from nsvf_b200.clib import _ext
dev = torch.device("cuda:0")
for name in ("C2", "C3", "C4"):
    scene = synthetic.make_scene(name)
    pts = torch.from_numpy(scene.points).to(dev)
    res = 1600 if name == "C2" else 800
    rs, rd = synthetic.camera_rays(res, res, 1, radius=4.5, seed=7, device=dev)
    rs = rs.expand_as(rd).reshape(1, -1, 3).contiguous(); rd = rd.reshape(1, -1, 3).contiguous()
    P = scene.max_hits
    out = {}
    for tag in ("walk", "tree"):
        if tag == "tree": os.environ["NSVF_AABB_NO_GRID"] = "1"
        else: os.environ.pop("NSVF_AABB_NO_GRID", None)
        ts = bench._time(lambda: _ext.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, P, 10000.0, shared_points=True), n=5, warm=2)
        ti = bench._time(lambda: _ext.aabb_intersect(rs, rd, pts, scene.voxel_size, P, shared_points=True), n=5, warm=2)
        ta = bench._time(lambda: _ext.aabb_hit_mask(rs, rd, pts, scene.voxel_size, shared_points=True), n=5, warm=2)
        out[tag] = (ts, ti, ta)
        res_s = _ext.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, P, 10000.0, shared_points=True)
        out[tag + "_r"] = res_s
    same = all(torch.equal(a, b) for a, b in zip(out["walk_r"], out["tree_r"]))
    hits = (out["walk_r"][0] >= 0).sum(-1).float()
    print("%s voxels %d rays %d P %d | sorted %.3f / %.3f ms, index %.3f / %.3f ms, any-hit %.3f / %.3f ms (walk / tree) | identical %s | hits/ray mean %.1f max %d"
          % (name, pts.shape[0], rs.shape[1], P, out["walk"][0], out["tree"][0], out["walk"][1], out["tree"][1], out["walk"][2], out["tree"][2], same, float(hits.mean()), int(hits.max())))
os.environ.pop("NSVF_AABB_NO_GRID", None)
