"""Synthetic scenes and rays of the shapes named in BASELINE.json (no datasets are available offline).

Everything here is plain numpy/torch host code used by tests/ and bench.py to build inputs; it follows
the reference's conventions so that the tensors look like what SparseVoxelEncoder would hold:
  * voxel centres on a regular grid in np.meshgrid default ('xy') order, fairnr/modules/encoder.py:1053-1059
    (bbox2voxels): voxel index = iy*(nx*nz) + ix*nz + iz;
  * 8 corner keys per voxel, corner order x slowest / z fastest (offset_points, geometry.py:229-238),
    keys numbered in lexicographic order of their integer coordinates (torch.unique(dim=0, sorted=True),
    encoder.py:272-275).
"""
import math

import numpy as np
import torch


def bbox_voxels(vmin, vmax, voxel_size):
    """Voxel centres of a bbox whose min/max are themselves centres (encoder.py:1053-1059)."""
    vmin, vmax = np.asarray(vmin, np.float64), np.asarray(vmax, np.float64)
    steps = np.round((vmax - vmin) / voxel_size).astype(np.int64) + 1
    x, y, z = [c.reshape(-1).astype("float32") for c in
               np.meshgrid(np.arange(steps[0]), np.arange(steps[1]), np.arange(steps[2]))]
    x, y, z = x * voxel_size + vmin[0], y * voxel_size + vmin[1], z * voxel_size + vmin[2]
    return np.stack([x, y, z]).T.astype("float32")


def carve_shell(points, r_in=0.6, r_out=1.0):
    """Keep voxels whose centre lies in a spherical shell (object-like sparsity), SURVEY.md §8d."""
    r = np.linalg.norm(points, axis=1)
    R = r.max() / math.sqrt(3.0) * 1.2
    keep = (r > r_in * R) & (r < r_out * R)
    return points[keep]


def corner_keys(points, voxel_size):
    """feats int64 [n,8], keys int64 [Kc,3]: the reference's discretize_points + offset_points + unique
    (encoder.py:270-275) restated with numpy integers."""
    half = voxel_size * 0.5
    pmin = points.min(0, keepdims=True)
    coords = np.round((points - pmin) / np.float32(half)).astype(np.int64)       # voxel centres: 2 apart
    off = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], np.int64)
    keys0 = (coords[:, None, :] + off[None]).reshape(-1, 3)
    keys, inv = np.unique(keys0, axis=0, return_inverse=True)
    return inv.reshape(-1, 8).astype(np.int64), keys


def split_points(points, voxel_size, times=1):
    """Children centres only (the geometric half of splitting_points, geometry.py:250-253)."""
    off = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], np.float32)
    for _ in range(times):
        quarter = np.float32(voxel_size * 0.25)
        points = (points[:, None, :] + off[None] * quarter).reshape(-1, 3)
        voxel_size = voxel_size * 0.5
    return points.astype("float32"), voxel_size


def camera_rays(height, width, n_views, radius=3.5, fov_focal=1111.0, seed=0, device="cpu", dtype=torch.float32):
    """Pinhole cameras on a sphere looking at the origin: ray_start [V,1,3], ray_dir [V,H*W,3] (unit norm)."""
    rng = np.random.RandomState(seed)
    starts, dirs = [], []
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64) + 0.5, np.arange(width, dtype=np.float64) + 0.5,
                         indexing="ij")
    for v in range(n_views):
        phi = rng.uniform(0, 2 * math.pi)
        theta = rng.uniform(math.radians(35), math.radians(85))
        eye = radius * np.array([math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)])
        fwd = -eye / np.linalg.norm(eye)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        d = (fwd[None, None] * fov_focal * (width / 800.0) + right[None, None] * (xs - width / 2)[..., None]
             + down[None, None] * (ys - height / 2)[..., None])
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        starts.append(eye[None])
        dirs.append(d.reshape(-1, 3))
    rs = torch.tensor(np.stack(starts), dtype=dtype, device=device)
    rd = torch.tensor(np.stack(dirs), dtype=dtype, device=device)
    return rs, rd


def random_rays(n, radius=3.0, target_extent=0.8, seed=0):
    """n rays from random points on a sphere towards random points near the origin (CPU numpy)."""
    rng = np.random.RandomState(seed)
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * radius
    t = rng.uniform(-target_extent, target_extent, size=(n, 3))
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype("float32"), d.astype("float32")


class Scene:
    """A voxel set with corner keys and random-init embeddings, as SparseVoxelEncoder buffers."""

    def __init__(self, points, voxel_size, embed_dim=32, max_hits=60, step_ratio=0.125, seed=0):
        self.points = np.ascontiguousarray(points, np.float32)
        self.voxel_size = float(voxel_size)
        self.step_size = float(step_ratio * voxel_size)
        self.max_hits = int(max_hits)
        self.feats, self.keys = corner_keys(self.points, voxel_size)
        rng = np.random.RandomState(seed)
        # Embedding init: normal(0, D^-0.5), fairnr/modules/module_utils.py:23-26
        self.values = (rng.normal(size=(len(self.keys), embed_dim)) * embed_dim ** -0.5).astype("float32")

    @property
    def n(self):
        return len(self.points)


def make_scene(name, seed=0):
    """Named configurations of BASELINE.json / BASELINE.md §3."""
    if name == "C1":      # CPU plumbing: bbox centres +-0.875, voxel 0.25 -> 512 voxels
        return Scene(bbox_voxels([-0.875] * 3, [0.875] * 3, 0.25), 0.25, max_hits=60, seed=seed)
    if name == "C2":      # nsvf_base training: bbox centres +-1.2, voxel 0.4 -> 343 voxels, step 0.05
        return Scene(bbox_voxels([-1.2] * 3, [1.2] * 3, 0.4), 0.4, max_hits=60, seed=seed)
    if name in ("C3", "C4"):   # 13^3 grid carved to 1752 voxels, split 2x (C3: 112 128 voxels) / 3x (C4: 897 024)
        pts = carve_shell(bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4), 0.35, 1.1)
        times = 2 if name == "C3" else 3
        pts, vs = split_points(pts, 0.4, times)
        return Scene(pts, vs, max_hits=135 if name == "C3" else 202, seed=seed)
    if name == "C5":      # Tanks&Temples-shaped elongated bbox (12 x 8 x 6.4), thin shell, 1 split -> 28 464 voxels of 0.2
        pts = carve_shell(bbox_voxels([-6.0, -4.0, -3.2], [6.0, 4.0, 3.2], 0.4), 0.75, 1.0)
        pts, vs = split_points(pts, 0.4, 1)
        return Scene(pts, vs, max_hits=90, seed=seed)
    raise ValueError(name)
