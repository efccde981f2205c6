"""Host-side mirror of the geometry helpers on the hot path (fairnr/data/geometry.py:195-327).

  offset_points      :229-238     corner / child / lattice offsets, x slowest - z fastest
  discretize_points  :241-247
  corner_keys        encoder.py:270-275 (discretize + offset + unique)
  build_easy_octree  :320-327     -> nsvf_octree_build (host, one D2H copy)
  splitting_points   :250-274     -> split kernels (see nsvf_b200/csrc/split.cu) with deterministic dedup
  trilinear_interp   :195-200     -> ops.trilinear_embed (fused gather kernel)
"""
import torch

from .clib import _ext


def offset_points(point_xyz, quarter_voxel=1, offset_only=False, bits=2):
    c = torch.arange(1, 2 * bits, 2, device=point_xyz.device)
    ox, oy, oz = torch.meshgrid([c, c, c], indexing="ij")
    offset = (torch.stack([ox.reshape(-1), oy.reshape(-1), oz.reshape(-1)], 1).type_as(point_xyz) - bits) / float(bits - 1)
    if not offset_only:
        return point_xyz.unsqueeze(1) + offset.unsqueeze(0).type_as(point_xyz) * quarter_voxel
    return offset.type_as(point_xyz) * quarter_voxel


def discretize_points(voxel_points, voxel_size):
    minimal_voxel_point = voxel_points.min(dim=0, keepdim=True)[0]
    voxel_indices = ((voxel_points - minimal_voxel_point) / voxel_size).round_().long()
    residual = (voxel_points - voxel_indices.type_as(voxel_points) * voxel_size).mean(0, keepdim=True)
    return voxel_indices, residual


def corner_keys(points, half_voxel):
    """(feats int64 [n,8], keys int64 [Kc,3]) — lexicographically sorted unique corner coordinates."""
    coords, _ = discretize_points(points, half_voxel)
    keys0 = offset_points(coords, 1.0).reshape(-1, 3)
    keys, feats = torch.unique(keys0, dim=0, sorted=True, return_inverse=True)
    return feats.reshape(-1, 8), keys


def build_easy_octree(points, half_voxel):
    coords, residual = discretize_points(points, half_voxel)
    ranges = coords.max(0)[0] - coords.min(0)[0]
    depths = torch.log2(ranges.max().float()).ceil_().long() - 1
    center = (coords.max(0)[0] + coords.min(0)[0]) / 2
    centers, children = _ext.build_octree(center, coords, int(depths))
    centers = centers.float() * half_voxel + residual
    return centers, children


def splitting_points(point_xyz, point_feats, values, half_voxel):
    """Half-voxel splitting: 8 children per voxel, deduplicated corner keys in lexicographic order,
    new embeddings interpolated inside the parent voxel.  Returns (new_points [8n,3], new_feats i64 [8n,8],
    new_values [Kc',D] or None, new_keys i64 [Kc',3]).  Parent choice per new key is the MINIMUM voxel index
    that touches it (the reference's scatter_ with duplicate indices is order-undefined on CUDA; the
    interpolated value is the same up to rounding because trilinear interpolation is continuous across faces)."""
    from . import split
    return split.splitting_points(point_xyz, point_feats, values, half_voxel)
