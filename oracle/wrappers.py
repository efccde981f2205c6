"""Restatement of the reference's PYTHON glue around the native kernels — TEST INFRASTRUCTURE ONLY.

Each function restates one piece of reference host code (cited) in torch, parametrised by a Level-1
backend `ext` that has the reference `_ext` signatures.  With `ext = oracle._ref ref_ext` (the
unmodified reference CUDA kernels) this reproduces the reference's Level-2 behaviour on the GPU box;
with `ext = NumpyExt()` (the C oracle) it runs on CPU.
"""
import math

import numpy as np
import torch

import oracle

MAX_DEPTH = 10000.0


class NumpyExt:
    """`_ext`-shaped adaptor over the C oracle (CPU tensors in / out)."""

    def __init__(self, inv_dir=None):
        self.inv_dir = inv_dir

    @staticmethod
    def _t(*arrs):
        return tuple(torch.from_numpy(a) for a in arrs)

    def aabb_intersect(self, ray_start, ray_dir, points, voxelsize, n_max):
        return self._t(*oracle.aabb_intersect(ray_start.numpy(), ray_dir.numpy(), points.numpy(), float(voxelsize),
                                              int(n_max), self.inv_dir))

    def svo_intersect(self, ray_start, ray_dir, points, children, voxelsize, n_max):
        return self._t(*oracle.svo_intersect(ray_start.numpy(), ray_dir.numpy(), points.numpy(), children.numpy(),
                                             float(voxelsize), int(n_max), self.inv_dir))

    def uniform_ray_sampling(self, pts_idx, min_depth, max_depth, noise, step_size, max_steps):
        return self._t(*oracle.uniform_ray_sampling(pts_idx.numpy(), min_depth.numpy(), max_depth.numpy(),
                                                    noise.numpy(), float(step_size), int(max_steps)))

    def inverse_cdf_sampling(self, pts_idx, min_depth, max_depth, noise, probs, steps, fixed_step_size):
        return self._t(*oracle.inverse_cdf_sampling(pts_idx.numpy(), min_depth.numpy(), max_depth.numpy(),
                                                    noise.numpy(), probs.numpy(), steps.numpy(),
                                                    float(fixed_step_size)))


def _tile_rays(ray_start, ray_dir, G):
    """The ray padding / tiling HACK of clib/__init__.py:63-70: wrap-pad N to K*G, view as [S*G, K, 3]."""
    S, N = ray_start.shape[:2]
    K = int(np.ceil(N / G))
    H = K * G
    if H > N:
        ray_start = torch.cat([ray_start, ray_start[:, :H - N]], 1)
        ray_dir = torch.cat([ray_dir, ray_dir[:, :H - N]], 1)
    return ray_start.reshape(S * G, K, 3), ray_dir.reshape(S * G, K, 3), S, N, H


def aabb_ray_intersect(ext, voxelsize, n_max, points, ray_start, ray_dir, G=None):
    """AABBRayIntersect.forward, clib/__init__.py:58-95 (G defaults to the reference's formula)."""
    if G is None:
        G = min(2048, int(2 * 10 ** 9 / points.numel()))
    rs, rd, S, N, H = _tile_rays(ray_start, ray_dir, G)
    pts = points.expand(S * G, *points.size()[1:]).contiguous()
    inds, dmin, dmax = ext.aabb_intersect(rs.float().contiguous(), rd.float().contiguous(), pts.float(), voxelsize, n_max)
    out = [t.reshape(S, H, -1)[:, :N] for t in (inds, dmin.type_as(ray_start), dmax.type_as(ray_start))]
    return tuple(out)


def svo_ray_intersect(ext, voxelsize, n_max, points, children, ray_start, ray_dir, G=None):
    """SparseVoxelOctreeRayIntersect.forward, clib/__init__.py:98-135."""
    if G is None:
        G = min(2048, int(2 * 10 ** 9 / (points.numel() + children.numel())))
    rs, rd, S, N, H = _tile_rays(ray_start, ray_dir, G)
    pts = points.expand(S * G, *points.size()[1:]).contiguous()
    ch = children.expand(S * G, *children.size()[1:]).contiguous()
    inds, dmin, dmax = ext.svo_intersect(rs.float().contiguous(), rd.float().contiguous(), pts.float(), ch.int(),
                                         voxelsize, n_max)
    out = [t.reshape(S, H, -1)[:, :N] for t in (inds, dmin.type_as(ray_start), dmax.type_as(ray_start))]
    return tuple(out)


def inverse_cdf_sampling(ext, pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False,
                         noise=None):
    """InverseCDFRaySampling.forward, clib/__init__.py:231-300.  `noise` (optional, [G, H/G, max_steps])
    replaces the wrapper's own draw so that two implementations can be fed identical noise."""
    G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
    H = int(np.ceil(N / G)) * G
    if H > N:
        pad = H - N
        pts_idx = torch.cat([pts_idx, pts_idx[:1].expand(pad, P)], 0)
        min_depth = torch.cat([min_depth, min_depth[:1].expand(pad, P)], 0)
        max_depth = torch.cat([max_depth, max_depth[:1].expand(pad, P)], 0)
        probs = torch.cat([probs, probs[:1].expand(pad, P)], 0)
        steps = torch.cat([steps, steps[:1].expand(pad)], 0)
    pts_idx, min_depth, max_depth, probs = [t.reshape(G, -1, P) for t in (pts_idx, min_depth, max_depth, probs)]
    steps = steps.reshape(G, -1)
    max_steps = int(steps.ceil().long().max()) + P
    if noise is None:
        noise = min_depth.new_zeros(*min_depth.size()[:-1], max_steps)
        noise = noise + 0.5 if deterministic else noise.uniform_().clamp(min=0.001, max=0.999)
    chunk = 4 * G
    parts = []
    for i in range(0, min_depth.size(1), chunk):
        sl = slice(i, i + chunk)
        parts.append(ext.inverse_cdf_sampling(
            pts_idx[:, sl].contiguous(), min_depth.float()[:, sl].contiguous(), max_depth.float()[:, sl].contiguous(),
            noise.float()[:, sl].contiguous(), probs.float()[:, sl].contiguous(), steps.float()[:, sl].contiguous(),
            fixed_step_size))
    sidx, sdepth, sdists = [torch.cat([p[k] for p in parts], 1).reshape(H, -1)[:N] for k in range(3)]
    max_len = int(sidx.ne(-1).sum(-1).max())
    return sidx[:, :max_len], sdepth[:, :max_len].type_as(min_depth), sdists[:, :max_len].type_as(min_depth)


def sort_hits(pts_idx, min_depth, max_depth):
    """SparseVoxelEncoder.ray_intersect post-processing, fairnr/modules/encoder.py:519-524."""
    min_depth = min_depth.masked_fill(pts_idx.eq(-1), MAX_DEPTH)
    max_depth = max_depth.masked_fill(pts_idx.eq(-1), MAX_DEPTH)
    min_depth, order = min_depth.sort(dim=-1)
    max_depth = max_depth.gather(-1, order)
    pts_idx = pts_idx.gather(-1, order)
    hits = pts_idx.ne(-1).any(-1)
    return pts_idx, min_depth, max_depth, hits


def probs_and_steps(pts_idx, min_depth, max_depth, step_size):
    """NSVFModel.intersecting, fairnr/models/nsvf.py:65-74."""
    dists = (max_depth - min_depth).masked_fill(pts_idx.eq(-1), 0)
    probs = dists / dists.sum(dim=-1, keepdim=True)
    steps = dists.sum(-1) / step_size
    return probs, steps


def mask_samples(sampled_idx, sampled_depth, sampled_dists):
    """SparseVoxelEncoder.ray_sample post-processing, fairnr/modules/encoder.py:547-549."""
    sampled_dists = sampled_dists.clamp(min=0.0)
    sampled_depth = sampled_depth.masked_fill(sampled_idx.eq(-1), MAX_DEPTH)
    sampled_dists = sampled_dists.masked_fill(sampled_idx.eq(-1), 0.0)
    return sampled_idx, sampled_depth, sampled_dists


def trilinear_torch(sampled_idx, sampled_xyz, feats, centres, values, voxel_size):
    """Plain-PyTorch statement of encoder.py:582-590 + geometry.py:195-200, 229-238 (differentiable)."""
    import torch.nn.functional as F
    c = F.embedding(sampled_idx, centres)
    e = F.embedding(F.embedding(sampled_idx, feats), values).view(c.size(0), 8, -1)
    p = ((sampled_xyz - c) / voxel_size + .5).unsqueeze(1)
    q = torch.tensor([[a, b, d] for a in (0., 1.) for b in (0., 1.) for d in (0., 1.)], dtype=p.dtype,
                     device=p.device).unsqueeze(0)
    w = (p * q + (1 - p) * (1 - q)).prod(dim=-1, keepdim=True)
    return (w * e).sum(1)


def composite_torch(free_energy, texture, sampled_depth):
    """Plain-PyTorch statement of renderer.py:193-218 (differentiable)."""
    shifted = torch.cat([free_energy.new_zeros(sampled_depth.size(0), 1), free_energy[:, :-1]], dim=-1)
    a = 1 - torch.exp(-free_energy.float())
    b = torch.exp(-torch.cumsum(shifted.float(), dim=-1))
    probs = (a * b).type_as(free_energy)
    depth = (sampled_depth * probs).sum(-1)
    missed = 1 - probs.sum(-1)
    colors = (texture * probs.unsqueeze(-1)).sum(-2)
    return probs, depth, missed, colors


def fill_in_blend_torch(hits, colors, missed, depths, bg_color, bg_depth):
    """fill_in (fairnr/data/geometry.py:303-317) for missed / colors / depths + the background blend of
    NSVFModel.postprocessing (fairnr/models/nsvf.py:93-104), plain torch."""
    n = hits.numel()
    full_m = missed.new_ones(n).masked_scatter(hits, missed)
    full_c = colors.new_zeros(n, 3).masked_scatter(hits.unsqueeze(-1).expand(n, 3), colors)
    full_d = depths.new_zeros(n).masked_scatter(hits, depths)
    full_c = full_c + full_m.unsqueeze(-1) * bg_color.reshape(1, 3)
    full_d = full_d + full_m * bg_depth
    return full_c, full_m, full_d


def track_voxel_probs_torch(max_voxel_probs, voxel_idxs, voxel_probs):
    """SparseVoxelEncoder.track_voxel_probs, fairnr/modules/encoder.py:594-603, plain torch."""
    n = max_voxel_probs.size(0)
    voxel_idxs = voxel_idxs.masked_fill(voxel_idxs.eq(-1), n)
    for start in range(0, voxel_idxs.size(0), 4096):
        end = min(start + 4096, voxel_idxs.size(0))
        cur = max_voxel_probs.new_zeros(end - start, n + 1).scatter_add_(
            dim=-1, index=voxel_idxs[start:end], src=voxel_probs[start:end]).max(0)[0][:-1]
        max_voxel_probs = torch.max(max_voxel_probs, cur)
    return max_voxel_probs


def uniform_ray_sampling(ext, pts_idx, min_depth, max_depth, step_size, max_ray_length, deterministic=False, noise=None):
    """UniformRaySampling.forward, clib/__init__.py:178-228 (wrap-padding to a multiple of 256 rows)."""
    G, N, P = 256, pts_idx.size(0), pts_idx.size(1)
    H = int(np.ceil(N / G)) * G
    if H > N:
        pts_idx = torch.cat([pts_idx, pts_idx[:H - N]], 0)
        min_depth = torch.cat([min_depth, min_depth[:H - N]], 0)
        max_depth = torch.cat([max_depth, max_depth[:H - N]], 0)
    pts_idx, min_depth, max_depth = [t.reshape(G, -1, P) for t in (pts_idx, min_depth, max_depth)]
    max_steps = int(max_ray_length / step_size) + P * 2
    if noise is None:
        noise = min_depth.new_zeros(*min_depth.size()[:-1], max_steps)
        noise = noise + 0.5 if deterministic else noise.uniform_()
    sidx, sdepth, sdists = ext.uniform_ray_sampling(pts_idx.contiguous(), min_depth.float().contiguous(),
                                                    max_depth.float().contiguous(), noise.float(), step_size, max_steps)
    sidx, sdepth, sdists = [t.reshape(H, -1)[:N] for t in (sidx, sdepth, sdists)]
    max_len = int(sidx.ne(-1).sum(-1).max())
    return sidx[:, :max_len], sdepth[:, :max_len], sdists[:, :max_len]
