#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference CUDA kernels (oracle/_ref/ref_ext.so, fairnr/clib rebuilt
for sm_100a) on a B200:   gpurun -- python tests/golden/make_gpu_golden.py   -> gpurun_out/gpu_*.npz, which are
then committed under tests/golden/.  They pin the CPU oracle (tests/test_oracle_golden.py) with the exact
MUFU-based reciprocals stored next to the outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref, wrappers  # noqa: E402
from nsvf_b200 import synthetic  # noqa: E402
from nsvf_b200.clib import _ext as ours  # noqa: E402
from tests import helpers  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
ref = build_ref.load()
assert ref is not None, "oracle/_ref/ref_ext.so missing"
n = lambda t: t.detach().cpu().numpy()

# ---- aabb: C1 scene, random + degenerate rays ------------------------------------------------------------
scene = synthetic.make_scene("C1")
pts = torch.from_numpy(scene.points).to(dev)
o, d = synthetic.random_rays(250, seed=9)
y0, z0 = -0.875 + 0.25 * 2, -0.875 + 0.25 * 5
extra = [((-3.0, y0, z0), (1.0, 0.0, 0.0)), ((-3.0, y0 + 0.125, z0), (1.0, 0.0, 0.0)),
         ((-3.0, y0 + 0.125, z0), (1.0, -0.0, -0.0)), ((3.0, y0, z0), (-1.0, 0.0, 0.0)),
         ((0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), ((-3.0, -3.0, -3.0), (1.0, 1.0, 1.0))]
o = np.concatenate([o, np.array([e[0] for e in extra], np.float32)])
d = np.concatenate([d, np.array([e[1] for e in extra], np.float32)])
rs, rd = torch.from_numpy(o).to(dev)[None].contiguous(), torch.from_numpy(d).to(dev)[None].contiguous()
inv = helpers.ref_rcp(rd)
for n_max in (12, 3):
    idx, dmin, dmax = ref.aabb_intersect(rs, rd, pts[None].contiguous(), scene.voxel_size, n_max)
    np.savez_compressed(os.path.join(OUT, "gpu_aabb_nmax%d.npz" % n_max), points=scene.points, voxel_size=scene.voxel_size,
                        n_max=n_max, ray_start=o, ray_dir=d, inv_dir=n(inv[0]), idx=n(idx[0]), min_depth=n(dmin[0]),
                        max_depth=n(dmax[0]))

# ---- svo: carved 13^3 grid, octree from the reference builder ------------------------------------------------
p0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
centers, children = helpers.easy_octree(torch.from_numpy(p0), 0.4, ref.build_octree)
centers, children = centers.to(dev).contiguous(), children.to(dev).contiguous()
o2, d2 = synthetic.random_rays(256, radius=4.5, target_extent=2.0, seed=4)
rs2, rd2 = torch.from_numpy(o2).to(dev)[None].contiguous(), torch.from_numpy(d2).to(dev)[None].contiguous()
inv2 = helpers.ref_rcp(rd2)
for n_max in (20, 2):
    idx, dmin, dmax = ref.svo_intersect(rs2, rd2, centers[None].contiguous(), children[None].contiguous(), 0.4, n_max)
    np.savez_compressed(os.path.join(OUT, "gpu_svo_nmax%d.npz" % n_max), centers=n(centers), children=n(children),
                        voxel_size=0.4, n_max=n_max, ray_start=o2, ray_dir=d2, inv_dir=n(inv2[0]), idx=n(idx[0]),
                        min_depth=n(dmin[0]), max_depth=n(dmax[0]))

# ---- sampling: inputs from the real pipeline ---------------------------------------------------------------
o3, d3 = synthetic.random_rays(400, seed=21)
idx, dmin, dmax = ref.aabb_intersect(torch.from_numpy(o3).to(dev)[None].contiguous(),
                                     torch.from_numpy(d3).to(dev)[None].contiguous(), pts[None].contiguous(),
                                     scene.voxel_size, 24)
idx, dmin, dmax, hits = wrappers.sort_hits(idx[0], dmin[0], dmax[0])
idx, dmin, dmax = idx[hits][:256], dmin[hits][:256], dmax[hits][:256]
probs, steps = wrappers.probs_and_steps(idx, dmin, dmax, scene.step_size)
G, R, P = 4, 64, 24
sh = lambda t: t.reshape(G, R, *t.shape[1:]).contiguous()
idx, dmin, dmax, probs, steps = map(sh, (idx, dmin, dmax, probs, steps))
max_steps = int(steps.ceil().max()) + P
torch.manual_seed(0)
noise = torch.zeros(G, R, max_steps, device=dev).uniform_().clamp(min=0.001, max=0.999)
for fixed in (-1.0, 0.02):
    si, sd, ss = ref.inverse_cdf_sampling(idx, dmin, dmax, noise, probs, steps, fixed)
    np.savez_compressed(os.path.join(OUT, "gpu_inverse_cdf_%s.npz" % ("auto" if fixed < 0 else "fixed")),
                        pts_idx=n(idx), min_depth=n(dmin), max_depth=n(dmax), noise=n(noise), probs=n(probs),
                        steps=n(steps), fixed_step_size=fixed, sampled_idx=n(si), sampled_depth=n(sd), sampled_dists=n(ss))
span = float((dmax.masked_fill(idx.eq(-1), 0).max(-1)[0] - dmin[..., 0]).max())
ms_u = int(span / scene.step_size) + 2 * P
noise_u = torch.zeros(G, R, ms_u, device=dev).uniform_().clamp(min=0.001, max=0.999)
si, sd, ss = ref.uniform_ray_sampling(idx, dmin, dmax, noise_u, scene.step_size, ms_u)
np.savez_compressed(os.path.join(OUT, "gpu_uniform.npz"), pts_idx=n(idx), min_depth=n(dmin), max_depth=n(dmax),
                    noise=n(noise_u), step_size=scene.step_size, max_steps=ms_u, sampled_idx=n(si), sampled_depth=n(sd),
                    sampled_dists=n(ss))
print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("gpu_")))
