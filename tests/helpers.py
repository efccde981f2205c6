"""Shared helpers for the parity tests (inputs are seeded; sizes finish in seconds on the oracle)."""
import numpy as np
import torch

from nsvf_b200 import _lib, synthetic

# Stated fp32 tolerance of the parity contract (BASELINE.json north_star): <= 1e-5 relative.
# Element-wise for values that are produced by the same op sequence; for sums whose order legitimately
# differs (atomics, parallel scans) the error is measured against the tensor's scale: max|a-b| <= RTOL*max|b|.
RTOL = 1e-5


def assert_close_scaled(a, b, rtol=RTOL, what=""):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a - b).abs().max())
    assert err <= rtol * scale, "%s: max abs err %.3e > %.1e * scale %.3e" % (what, err, rtol, scale)


def ref_rcp(ray_dir):
    """1/dir with the reference's intrinsic (__fdividef), computed on the GPU: fed to the CPU oracle."""
    L = _lib.load()
    out = torch.empty_like(ray_dir)
    _lib.check(L.nsvf_ref_rcp(_lib.current_stream(ray_dir.device), ray_dir.numel(), ray_dir.data_ptr(), out.data_ptr()))
    return out


def scene_tensors(scene, device):
    pts = torch.from_numpy(scene.points).to(device)
    feats = torch.from_numpy(scene.feats).to(device)
    values = torch.from_numpy(scene.values).to(device)
    return pts, feats, values


def rays_for(scene_name, n, seed, device):
    extent = {"C1": 0.8, "C2": 1.1}.get(scene_name, 2.0)
    o, d = synthetic.random_rays(n, radius=3.0 if scene_name in ("C1", "C2") else 4.5, target_extent=extent, seed=seed)
    return torch.from_numpy(o).to(device), torch.from_numpy(d).to(device)


def easy_octree(points, voxel_size, build_octree):
    """build_easy_octree, fairnr/data/geometry.py:320-327, on top of a Level-1 build_octree."""
    half = voxel_size / 2.0
    pmin = points.min(dim=0, keepdim=True)[0]
    coords = ((points - pmin) / half).round_().long()
    residual = (points - coords.type_as(points) * half).mean(0, keepdim=True)
    ranges = coords.max(0)[0] - coords.min(0)[0]
    depth = int(torch.log2(ranges.max().float()).ceil_().long() - 1)
    center = (coords.max(0)[0] + coords.min(0)[0]) / 2
    centers, children = build_octree(center, coords, depth)
    return centers.float() * half + residual, children
