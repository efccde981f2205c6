"""Level-2 boundary: the callables of the reference's `fairnr/clib/__init__.py` (same names, argument
order and results), running on the sm_100a kernels.

    aabb_ray_intersect(voxelsize, n_max, points, ray_start, ray_dir)              clib/__init__.py:58-95
    svo_ray_intersect(voxelsize, n_max, points, children, ray_start, ray_dir)     :98-135
    uniform_ray_sampling(pts_idx, min_depth, max_depth, step_size, max_ray_length, deterministic)  :178-228
    inverse_cdf_sampling(pts_idx, min_depth, max_depth, probs, steps, fixed_step_size, deterministic)  :231-300

What is NOT carried over from the reference wrappers (results are unchanged, see tests/test_clib_gpu.py):
  * the voxel set / octree is never replicated G <= 2048 times (`points.expand(S*G, ...).contiguous()`,
    up to 8 GB) and rays are not padded and re-tiled: our kernels take [S, N, 3] rays directly;
  * the inverse-CDF wrapper's padding rows (copies of ray 0), its 800-column chunk loop with six
    .contiguous() slices per chunk, and the full-tensor `ne(-1).sum(-1).max()` reduction are folded
    into one kernel launch (`valid_rays`, `ray_chunk`, `max_count` of nsvf_inverse_cdf_sampling);
    with `deterministic=True` the constant 0.5 noise tensor is not materialised at all.
  * RNG contract kept: non-deterministic noise is drawn exactly like the reference does
    (`new_zeros(G, H/G, max_steps).uniform_().clamp(0.001, 0.999)`), so a seeded run reproduces the
    reference's samples.
All outputs are marked non-differentiable and backward returns None, as in the reference.
"""
import numpy as np
import torch
from torch.autograd import Function

from .. import _lib
from . import _ext

MAX_DEPTH = 10000.0
_L = _lib.load()
_p = _lib.ptr


class AABBRayIntersect(Function):
    @staticmethod
    def forward(ctx, voxelsize, n_max, points, ray_start, ray_dir):
        inds, min_depth, max_depth = _ext.aabb_intersect(
            ray_start.float().contiguous(), ray_dir.float().contiguous(), points.float().contiguous(),
            voxelsize, n_max)
        min_depth = min_depth.type_as(ray_start)
        max_depth = max_depth.type_as(ray_start)
        ctx.mark_non_differentiable(inds)
        ctx.mark_non_differentiable(min_depth)
        ctx.mark_non_differentiable(max_depth)
        return inds, min_depth, max_depth

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None


aabb_ray_intersect = AABBRayIntersect.apply


class SparseVoxelOctreeRayIntersect(Function):
    @staticmethod
    def forward(ctx, voxelsize, n_max, points, children, ray_start, ray_dir):
        inds, min_depth, max_depth = _ext.svo_intersect(
            ray_start.float().contiguous(), ray_dir.float().contiguous(), points.float().contiguous(),
            children.int().contiguous(), voxelsize, n_max)
        min_depth = min_depth.type_as(ray_start)
        max_depth = max_depth.type_as(ray_start)
        ctx.mark_non_differentiable(inds)
        ctx.mark_non_differentiable(min_depth)
        ctx.mark_non_differentiable(max_depth)
        return inds, min_depth, max_depth

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None, None


svo_ray_intersect = SparseVoxelOctreeRayIntersect.apply


class UniformRaySampling(Function):
    """fairnr/clib/__init__.py:178-228.  The reference pads the rays to a multiple of 256 (wrapping), tiles them
    [256, R, P], runs its kernel and trims to the longest ray with a full-tensor reduction; our kernel takes the [N, P]
    rays as they are and reports the longest ray itself.  RNG contract kept: the noise is drawn with the numel and flat
    order of the reference's padded [256, R, max_steps] tensor, the first N rows are used."""

    @staticmethod
    def forward(ctx, pts_idx, min_depth, max_depth, step_size, max_ray_length, deterministic=False):
        N, P = pts_idx.size(0), pts_idx.size(1)
        dev = pts_idx.device
        step_size = float(step_size)
        max_steps = int(max_ray_length / step_size) + P * 2
        if deterministic:
            noise = min_depth.new_full((N, max_steps), 0.5, dtype=torch.float32)
        else:
            H = int(np.ceil(N / 256)) * 256
            noise = min_depth.new_zeros(H, max_steps, dtype=torch.float32).uniform_()[:N]
        idx32 = pts_idx.int().contiguous()
        dmin, dmax = min_depth.float().contiguous(), max_depth.float().contiguous()
        sampled_idx = torch.empty((N, max_steps), dtype=torch.int32, device=dev)
        sampled_depth = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
        sampled_dists = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
        max_count = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_uniform_ray_sampling(
                _lib.current_stream(dev), 1, N, P, max_steps, step_size, _p(idx32), _p(dmin), _p(dmax), _p(noise),
                _p(sampled_idx), _p(sampled_depth), _p(sampled_dists), _p(max_count)))
        max_len = int(max_count.item())
        sampled_idx = sampled_idx[:, :max_len]
        sampled_depth = sampled_depth[:, :max_len].type_as(min_depth)
        sampled_dists = sampled_dists[:, :max_len].type_as(min_depth)
        ctx.mark_non_differentiable(sampled_idx)
        ctx.mark_non_differentiable(sampled_depth)
        ctx.mark_non_differentiable(sampled_dists)
        return sampled_idx, sampled_depth, sampled_dists

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None, None


uniform_ray_sampling = UniformRaySampling.apply


class InverseCDFRaySampling(Function):
    @staticmethod
    def forward(ctx, pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False):
        G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
        R = int(np.ceil(N / G))          # rays per block row of the reference's [G, R, P] tiling
        dev = pts_idx.device
        in_dtype = min_depth.dtype

        pts_idx = pts_idx.int().contiguous()
        min_depth = min_depth.float().contiguous()
        max_depth = max_depth.float().contiguous()
        probs = probs.float().contiguous()
        steps = steps.float().contiguous()

        # reference: max_steps = steps.ceil().long().max() + P over the padded rays (copies of ray 0)
        max_steps = int(steps.ceil().long().max()) + P
        if deterministic:
            noise, noise_ptr = None, None
        else:
            # identical draw to the reference: numel G*R*max_steps from the current generator
            noise = min_depth.new_zeros(G, R, max_steps).uniform_().clamp(min=0.001, max=0.999)
            noise_ptr = _p(noise)

        sampled_idx = torch.empty((N, max_steps), dtype=torch.int32, device=dev)
        sampled_depth = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
        sampled_dists = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
        max_count = torch.zeros(1, dtype=torch.int32, device=dev)
        fixed = float(fixed_step_size)
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_inverse_cdf_sampling(
                _lib.current_stream(dev), G, R, N, 4 * G, P, max_steps, fixed, _p(pts_idx), _p(min_depth),
                _p(max_depth), noise_ptr, 0.5, _p(probs), _p(steps), _p(sampled_idx), _p(sampled_depth),
                _p(sampled_dists), _p(max_count)))
        sampled_depth = sampled_depth.to(in_dtype)
        sampled_dists = sampled_dists.to(in_dtype)

        max_len = int(max_count.item())
        sampled_idx = sampled_idx[:, :max_len]
        sampled_depth = sampled_depth[:, :max_len]
        sampled_dists = sampled_dists[:, :max_len]

        ctx.mark_non_differentiable(sampled_idx)
        ctx.mark_non_differentiable(sampled_depth)
        ctx.mark_non_differentiable(sampled_dists)
        return sampled_idx, sampled_depth, sampled_dists

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None, None, None


inverse_cdf_sampling = InverseCDFRaySampling.apply


@torch.no_grad()
def inverse_cdf_sampling_rows(pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False,
                              trimmed=False, pad_depth=MAX_DEPTH):
    """inverse_cdf_sampling + the post-processing of SparseVoxelEncoder.ray_sample (encoder.py:547-549: dists clamped at
    0, depth = MAX_DEPTH and dists = 0 where idx == -1) in ONE kernel, for the renderer's trimmed-row path:

        -> (sampled_idx i32, sampled_depth, sampled_dists [N, K], ray_len i32 [N], holes i32 [1])

    ray_len[r] = number of leading slots of row r that hold samples.  trimmed=False: K = max_len like the reference
    wrapper (one extra host sync) and the rows are padded (-1 / MAX_DEPTH / 0); trimmed=True: K = max_steps, nothing
    beyond ray_len[r] is written (or may be read), no max_len sync.  Same tiling quirks and RNG draw as
    InverseCDFRaySampling above."""
    G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
    R = int(np.ceil(N / G))
    dev = pts_idx.device
    in_dtype = min_depth.dtype
    pts_idx = pts_idx.int().contiguous()
    min_depth, max_depth = min_depth.float().contiguous(), max_depth.float().contiguous()
    probs, steps = probs.float().contiguous(), steps.float().contiguous()
    max_steps = int(steps.ceil().long().max()) + P
    if deterministic:
        noise, noise_ptr = None, None
    else:
        noise = min_depth.new_zeros(G, R, max_steps).uniform_().clamp(min=0.001, max=0.999)
        noise_ptr = _p(noise)
    sampled_idx = torch.empty((N, max_steps), dtype=torch.int32, device=dev)
    sampled_depth = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
    sampled_dists = torch.empty((N, max_steps), dtype=torch.float32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)          # [max_count, holes]
    ray_len = torch.empty(N, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_inverse_cdf_sampling_ex(
            _lib.current_stream(dev), G, R, N, 4 * G, P, max_steps, float(fixed_step_size), _p(pts_idx), _p(min_depth),
            _p(max_depth), noise_ptr, 0.5, _p(probs), _p(steps), _p(sampled_idx), _p(sampled_depth),
            _p(sampled_dists), _p(meta[0:1]), _p(ray_len), _p(meta[1:2]), float(pad_depth), (0 if trimmed else 1) | 2))
    if in_dtype != torch.float32:
        sampled_depth, sampled_dists = sampled_depth.to(in_dtype), sampled_dists.to(in_dtype)
    if not trimmed:
        max_len = int(meta[0].item())
        sampled_idx, sampled_depth, sampled_dists = (sampled_idx[:, :max_len], sampled_depth[:, :max_len],
                                                     sampled_dists[:, :max_len])
    return sampled_idx, sampled_depth, sampled_dists, ray_len, meta[1:2]


@torch.no_grad()
def inverse_cdf_sampling_lazy(pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False,
                              pad_depth=MAX_DEPTH):
    """On-demand sampling for the ray-marching plan: computes only how many samples each ray has (one kernel, no sample
    tensors) and returns what nsvf_inverse_cdf_block needs to materialise column blocks later, for the rays that are still
    alive.  Same tiling quirks / RNG draw as InverseCDFRaySampling.  -> dict of row-sliceable tensors + scalars."""
    G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
    R = int(np.ceil(N / G))
    dev = pts_idx.device
    pts_idx = pts_idx.int().contiguous()
    min_depth, max_depth = min_depth.float().contiguous(), max_depth.float().contiguous()
    probs, steps = probs.float().contiguous(), steps.float().contiguous()
    max_steps = int(steps.ceil().long().max()) + P
    noise = None if deterministic else min_depth.new_zeros(G, R, max_steps).uniform_().clamp(min=0.001, max=0.999)
    ray_len = torch.empty(N, dtype=torch.int32, device=dev)
    quirk = torch.empty((N, 2), dtype=torch.int32, device=dev)
    meta = torch.zeros(3, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_inverse_cdf_plan(
            _lib.current_stream(dev), G, R, N, 4 * G, P, max_steps, float(fixed_step_size), _p(pts_idx), _p(min_depth),
            _p(max_depth), _p(noise), 0.5, _p(probs), _p(steps), _p(ray_len), _p(quirk), _p(meta)))
    out = {"sampled_point_count": ray_len, "lazy_quirk": quirk, "lazy_pts_idx": pts_idx, "lazy_min_depth": min_depth,
           "lazy_max_depth": max_depth, "lazy_probs": probs, "lazy_steps": steps, "lazy_meta": meta,
           "lazy_max_steps": max_steps, "lazy_fixed_step_size": float(fixed_step_size), "lazy_pad_depth": float(pad_depth)}
    if noise is not None:
        out["lazy_noise"] = noise.view(G * R, max_steps)[:N]
    return out


class BallRayIntersect(Function):
    """fairnr/clib/__init__.py:38-55."""

    @staticmethod
    def forward(ctx, radius, n_max, points, ray_start, ray_dir):
        inds, min_depth, max_depth = _ext.ball_intersect(
            ray_start.float().contiguous(), ray_dir.float().contiguous(), points.float().contiguous(), radius, n_max)
        min_depth, max_depth = min_depth.type_as(ray_start), max_depth.type_as(ray_start)
        ctx.mark_non_differentiable(inds, min_depth, max_depth)
        return inds, min_depth, max_depth

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None


ball_ray_intersect = BallRayIntersect.apply


class TriangleRayIntersect(Function):
    """fairnr/clib/__init__.py:138-175 (no face replication: the face set is shared by all rays of a shape)."""

    @staticmethod
    def forward(ctx, cagesize, blur_ratio, n_max, points, faces, ray_start, ray_dir):
        import torch.nn.functional as F
        S, N = ray_start.shape[:2]
        face_points = F.embedding(faces.reshape(-1, 3), points.reshape(-1, 3)).reshape(1, -1, 9)
        face_points = face_points.expand(S, -1, -1).float().contiguous()
        inds, depth, uv = _ext.triangle_intersect(ray_start.float().contiguous(), ray_dir.float().contiguous(),
                                                  face_points, cagesize, blur_ratio, n_max)
        depth, uv = depth.type_as(ray_start), uv.type_as(ray_start)
        depth = depth.reshape(S, N, -1, 3)
        ctx.mark_non_differentiable(inds, depth, uv)
        return inds, depth, uv

    @staticmethod
    def backward(ctx, a, b, c):
        return None, None, None, None, None, None, None


triangle_ray_intersect = TriangleRayIntersect.apply
