"""Times the reference's own clib CUDA kernels (oracle/_ref/ref_ext.so = unmodified fairnr/clib rebuilt for sm_100a,
driven through a restatement of the reference wrappers) next to ours on the same B200 and the same tensors.
Not a pytest file:  gpurun -- python tests/perf/ref_clib_compare.py  -> gpurun_out/ref_clib_compare.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from nsvf_b200 import synthetic, clib, ops
from nsvf_b200.clib import _ext
from nsvf_b200 import geometry
from oracle import wrappers, build_ref

dev = torch.device("cuda:0")
ref = build_ref.load()
assert ref is not None


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)


out = {}
for name in ("C2", "C3"):
    scene = synthetic.make_scene(name)
    pts = torch.from_numpy(scene.points).to(dev)
    pts[:, 0] += scene.voxel_size / 10
    V = 4 if name == "C2" else 1
    rs, rd = synthetic.camera_rays(800, 800, V, radius=3.2 if name == "C2" else 4.5, seed=0, device=dev)
    rs = rs.expand_as(rd).reshape(1, -1, 3).contiguous()
    rd = rd.reshape(1, -1, 3).contiguous()
    P, vs = scene.max_hits, scene.voxel_size
    r = {"voxels": scene.n, "rays": rs.shape[1], "max_hits": P}
    n_ref = 2 if name == "C3" else 5
    r["aabb_ref_wrapper_ms"] = timeit(lambda: wrappers.aabb_ray_intersect(ref, vs, P, pts[None], rs, rd), n=n_ref, warm=1)
    r["aabb_ref_wrapper+sort_ms"] = timeit(lambda: wrappers.sort_hits(*wrappers.aabb_ray_intersect(ref, vs, P, pts[None], rs, rd)), n=n_ref, warm=1)
    r["aabb_ours_level2_ms"] = timeit(lambda: clib.aabb_ray_intersect(vs, P, pts[None], rs, rd))
    r["aabb_ours_sorted_fused_ms"] = timeit(lambda: _ext.aabb_intersect_sorted(rs, rd, pts, vs, P, 1e4, shared_points=True))
    r["aabb_ours_hit_mask_ms"] = timeit(lambda: _ext.aabb_hit_mask(rs, rd, pts, vs, shared_points=True))
    # octree variant
    centers, children = geometry.build_easy_octree(torch.from_numpy(scene.points).to(dev), vs / 2.0)
    centers, children = centers.contiguous(), children.contiguous()
    r["octree_nodes"] = centers.shape[0]
    r["svo_ref_wrapper_ms"] = timeit(lambda: wrappers.svo_ray_intersect(ref, vs, P, centers[None], children[None], rs, rd), n=n_ref, warm=1)
    r["svo_ours_level2_ms"] = timeit(lambda: clib.svo_ray_intersect(vs, P, centers[None], children[None], rs, rd))
    # sampling
    idx, dmin, dmax, hits = _ext.aabb_intersect_sorted(rs, rd, pts, vs, P, 1e4, shared_points=True)
    sel = hits[0].nonzero()[:, 0]
    if name == "C2":
        sel = sel[torch.randperm(sel.numel(), device=dev)[:8192]].sort()[0]
    idx, dmin, dmax = idx[0][sel].contiguous(), dmin[0][sel].contiguous(), dmax[0][sel].contiguous()
    probs, steps = wrappers.probs_and_steps(idx, dmin, dmax, scene.step_size)
    r["rays_marched"] = int(sel.numel())
    det = name != "C2"
    r["inverse_cdf_ref_wrapper_ms"] = timeit(lambda: wrappers.inverse_cdf_sampling(ref, idx, dmin, dmax, probs, steps, -1, det), n=n_ref, warm=1)
    r["inverse_cdf_ours_level2_ms"] = timeit(lambda: clib.inverse_cdf_sampling(idx, dmin, dmax, probs, steps, -1, det))
    sidx, sdep, sdist = wrappers.mask_samples(*clib.inverse_cdf_sampling(idx, dmin, dmax, probs, steps, -1, det))
    r["max_len"] = sidx.shape[1]
    # torch-level stages: reference statement (plain torch ops, as the reference runs them) vs fused kernels
    keep = slice(0, min(sidx.shape[0], 65536))
    sidx, sdep, sdist = sidx[keep], sdep[keep], sdist[keep]
    mask = sidx.ne(-1)
    o, d = rs[0][sel][keep], rd[0][sel][keep]
    xyz = (o[:, None] + d[:, None] * sdep[..., None])[mask]
    vox = sidx[mask].long()
    feats = torch.from_numpy(scene.feats).to(dev)
    values = torch.from_numpy(scene.values).to(dev)
    r["samples_interp"] = int(vox.numel())
    r["interp_fwd_ref_torch_ms"] = timeit(lambda: wrappers.trilinear_torch(vox, xyz, feats, pts, values, vs))
    r["interp_fwd_ours_ms"] = timeit(lambda: ops.trilinear_embed(vox, xyz, feats, pts, values, vs))
    v1, v2 = values.clone().requires_grad_(True), values.clone().requires_grad_(True)
    g = torch.randn(vox.numel(), 32, device=dev)
    r["interp_fwd+bwd_ref_torch_ms"] = timeit(lambda: wrappers.trilinear_torch(vox, xyz, feats, pts, v1, vs).backward(g))
    r["interp_fwd+bwd_ours_ms"] = timeit(lambda: ops.trilinear_embed(vox, xyz, feats, pts, v2, vs).backward(g))
    fe = (torch.rand_like(sdep) * 0.3 * mask).requires_grad_(True)
    tex = torch.rand(*sdep.shape, 3, device=dev).requires_grad_(True)
    r["composite_fwd_ref_torch_ms"] = timeit(lambda: wrappers.composite_torch(fe, tex, sdep))
    r["composite_fwd_ours_ms"] = timeit(lambda: ops.composite(fe, tex, sdep))
    r["composite_fwd+bwd_ref_torch_ms"] = timeit(lambda: sum(t.sum() for t in wrappers.composite_torch(fe, tex, sdep)[1:]).backward())
    r["composite_fwd+bwd_ours_ms"] = timeit(lambda: sum(t.sum() for t in ops.composite(fe, tex, sdep)[1:]).backward())
    ref_clib = r["aabb_ref_wrapper+sort_ms"] + r["inverse_cdf_ref_wrapper_ms"]
    ours_clib = r["aabb_ours_sorted_fused_ms"] + r["inverse_cdf_ours_level2_ms"]
    r["clib_rays_per_s_ref(intersect+sort+sample)"] = round(r["rays"] / ref_clib * 1e3)
    r["clib_rays_per_s_ours"] = round(r["rays"] / ours_clib * 1e3)
    r["clib_speedup"] = round(ref_clib / ours_clib, 2)
    out[name] = r
    print(name, json.dumps(r), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_clib_compare.json"), "w"), indent=1)
