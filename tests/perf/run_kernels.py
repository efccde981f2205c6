"""Launch each hand-written kernel once at bench sizes (for `ncu`; not a pytest file).
usage: python tests/perf/run_kernels.py [C2|C3] [kernel ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nsvf_b200 import synthetic, clib, ops
from nsvf_b200.clib import _ext
from oracle import wrappers

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
which = set(sys.argv[2:])
dev = torch.device("cuda:0")
scene = synthetic.make_scene(name)
pts = torch.from_numpy(scene.points).to(dev)
pts[:, 0] += scene.voxel_size / 10
feats = torch.from_numpy(scene.feats).int().to(dev)
values = torch.from_numpy(scene.values).to(dev)
V = 4 if name == "C2" else 1
rs, rd = synthetic.camera_rays(800, 800, V, radius=3.2 if name == "C2" else 4.5, seed=0, device=dev)
rs = rs.expand_as(rd).reshape(1, -1, 3).contiguous()
rd = rd.reshape(1, -1, 3).contiguous()
P = scene.max_hits
for rep in range(2):
    if not which or "aabb" in which:
        idx, dmin, dmax = _ext.aabb_intersect(rs, rd, pts, scene.voxel_size, P, shared_points=True)
    if not which or "aabb" in which:
        _ext.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, P, 1e4, shared_points=True)
        _ext.aabb_hit_mask(rs, rd, pts, scene.voxel_size, shared_points=True)
from nsvf_b200 import geometry
centers, children = geometry.build_easy_octree(torch.from_numpy(scene.points).to(dev), scene.voxel_size / 2.0)
for rep in range(2):
    if not which or "svo" in which:
        _ext.svo_intersect(rs, rd, centers.contiguous(), children.contiguous(), scene.voxel_size, P, shared_tree=True)
torch.cuda.synchronize()
idx, dmin, dmax = _ext.aabb_intersect(rs, rd, pts, scene.voxel_size, P, shared_points=True)
idx, dmin, dmax, hits = wrappers.sort_hits(idx[0], dmin[0], dmax[0])
n_keep = 8192 if name == "C2" else int(hits.sum())
sel = hits.nonzero()[:n_keep, 0]
idx, dmin, dmax = idx[sel].contiguous(), dmin[sel].contiguous(), dmax[sel].contiguous()
probs, steps = wrappers.probs_and_steps(idx, dmin, dmax, scene.step_size)
for rep in range(2):
    sidx, sdep, sdist = clib.inverse_cdf_sampling(idx, dmin, dmax, probs, steps, -1, name != "C2")
sidx, sdep, sdist = wrappers.mask_samples(sidx, sdep, sdist)
mask = sidx.ne(-1)
o, d = rs[0][sel], rd[0][sel]
xyz = (o[:, None] + d[:, None] * sdep[..., None])[mask]
vox = sidx[mask]
print(name, "voxels", scene.n, "rays", rs.shape[1], "marched", idx.shape[0], "K", sidx.shape[1], "samples", int(mask.sum()))
v2 = values.clone().requires_grad_(True)
for rep in range(2):
    emb = ops.trilinear_embed(vox, xyz, feats, pts, v2, scene.voxel_size)
    emb.backward(torch.ones_like(emb))
fe = (torch.rand_like(sdep) * 0.2 * mask).requires_grad_(True)
tex = torch.rand(*sdep.shape, 3, device=dev).requires_grad_(True)
for rep in range(2):
    out = ops.composite(fe, tex, sdep)
    (out[1].sum() + out[2].sum() + out[3].sum()).backward()
torch.cuda.synchronize()
print("done")
