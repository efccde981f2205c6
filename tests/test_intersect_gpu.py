"""GPU parity of ray/voxel intersection: ours (C ABI) vs the reference CUDA kernels (oracle/_ref) vs the
CPU oracle.  Integer outputs bit-exact; depths are produced by the identical fp32 op sequence, so they
are compared for equality too (torch.equal treats -0.0 == +0.0)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import wrappers
from nsvf_b200 import synthetic
from nsvf_b200 import clib
from nsvf_b200.clib import _ext as ours
from tests import helpers

pytestmark = pytest.mark.gpu


def _cmp3(a, b, what):
    for x, y, nm in zip(a, b, ("idx", "min_depth", "max_depth")):
        x, y = torch.as_tensor(x).cpu(), torch.as_tensor(y).cpu()
        assert x.shape == y.shape, (what, nm, x.shape, y.shape)
        assert torch.equal(x, y), "%s: %s differs in %d of %d entries" % (what, nm, int((x != y).sum()), x.numel())


def _oracle_aabb(rs, rd, pts, vs, n_max):
    inv = helpers.ref_rcp(rd)
    return oracle.aabb_intersect(rs.cpu().numpy(), rd.cpu().numpy(), pts.cpu().numpy(), vs, n_max, inv.cpu().numpy())


@pytest.mark.parametrize("name,n_rays,n_max", [("C1", 8192, 60), ("C1", 4096, 5), ("C2", 8192, 60)])
def test_aabb_level1_vs_reference_and_oracle(cuda, ref_ext, name, n_rays, n_max):
    scene = synthetic.make_scene(name)
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    o, d = helpers.rays_for(name, n_rays, 1, cuda)
    B = 4
    rs, rd = o.view(B, -1, 3).contiguous(), d.view(B, -1, 3).contiguous()
    ptsB = pts.unsqueeze(0).expand(B, -1, -1).contiguous()
    mine = ours.aabb_intersect(rs, rd, ptsB, scene.voxel_size, n_max)
    ref = ref_ext.aabb_intersect(rs, rd, ptsB, scene.voxel_size, n_max)
    _cmp3(mine, ref, "ours vs reference CUDA")
    _cmp3(mine, _oracle_aabb(rs, rd, ptsB, scene.voxel_size, n_max), "ours vs CPU oracle")
    assert int((mine[0] >= 0).sum()) > n_rays  # the case is not vacuous
    # shared voxel set (no replication) gives the same answer
    _cmp3(ours.aabb_intersect(rs, rd, pts, scene.voxel_size, n_max, shared_points=True), mine, "shared vs replicated")


def test_aabb_multilevel_hierarchy_vs_reference(cuda, ref_ext):
    """~14k and ~112k voxels: 3- and 4-level hierarchies, upper levels staged by TMA, level 0/1 from global."""
    for times, n_max, n_rays in ((1, 90, 4096), (2, 135, 2048)):
        pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
        p, vs = synthetic.split_points(pts0, 0.4, times)
        p[:, 0] += np.float32(vs / 10)  # the encoder's precompute shift (encoder.py:383)
        pts = torch.from_numpy(p).to(cuda)
        rs, rd = synthetic.camera_rays(64, 64, 1, device=cuda)
        rs = rs.expand_as(rd)[:, :n_rays].contiguous()
        rd = rd[:, :n_rays].contiguous()
        mine = ours.aabb_intersect(rs, rd, pts.unsqueeze(0), vs, n_max)
        ref = ref_ext.aabb_intersect(rs, rd, pts.unsqueeze(0).contiguous(), vs, n_max)
        _cmp3(mine, ref, "ours vs reference CUDA, n=%d" % len(p))
        assert int((mine[0] >= 0).sum(-1).max()) > 20


def test_aabb_truncation_keeps_lowest_indices(cuda, ref_ext):
    scene = synthetic.make_scene("C1")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    o, d = helpers.rays_for("C1", 2048, 3, cuda)
    full = ours.aabb_intersect(o[None], d[None], pts[None], scene.voxel_size, 60)[0][0]
    for n_max in (1, 2, 7, 33):
        cut = ours.aabb_intersect(o[None], d[None], pts[None], scene.voxel_size, n_max)[0][0]
        assert torch.equal(cut, full[:, :n_max])
        _cmp3(ours.aabb_intersect(o[None], d[None], pts[None], scene.voxel_size, n_max),
              ref_ext.aabb_intersect(o[None].contiguous(), d[None].contiguous(), pts[None].contiguous(),
                                     scene.voxel_size, n_max), "n_max=%d" % n_max)


def test_aabb_axis_aligned_and_degenerate_rays(cuda, ref_ext):
    """KAT-3 plus the NaN paths of the slab test: zero / negative-zero direction components, origins
    exactly on voxel face planes, origins inside voxels."""
    scene = synthetic.make_scene("C1")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    y0, z0 = -0.875 + 0.25 * 2, -0.875 + 0.25 * 5
    rays = [
        ((-3.0, y0, z0), (1.0, 0.0, 0.0)),
        ((-3.0, y0, z0), (1.0, -0.0, 0.0)),
        ((-3.0, y0 + 0.125, z0), (1.0, 0.0, 0.0)),        # on a face plane, +0 direction
        ((-3.0, y0 + 0.125, z0), (1.0, -0.0, -0.0)),      # on a face plane, -0 direction
        ((-3.0, y0 - 0.125, z0 + 0.125), (1.0, 0.0, -0.0)),
        ((3.0, y0, z0), (-1.0, 0.0, 0.0)),
        ((y0, 3.0, z0), (0.0, -1.0, 0.0)),
        ((y0, z0, -3.0), (-0.0, 0.0, 1.0)),
        ((0.0, 0.0, 0.0), (0.0, 0.0, 1.0)),               # starts inside, on lattice planes
        ((0.125, 0.125, 0.125), (1.0, 1.0, 1.0)),
        ((-3.0, -3.0, -3.0), (1.0, 1.0, 1.0)),            # exactly through corners
        ((-3.0, y0, z0), (1.0, 1e-30, -1e-30)),
        ((-3.0, y0, z0), (-1.0, 0.0, 0.0)),               # pointing away
        ((-3.0, y0, z0), (float("nan"), 0.0, 0.0)),
    ]
    rs = torch.tensor([r[0] for r in rays], dtype=torch.float32, device=cuda)[None].contiguous()
    rd = torch.tensor([r[1] for r in rays], dtype=torch.float32, device=cuda)[None].contiguous()
    mine = ours.aabb_intersect(rs, rd, pts[None].contiguous(), scene.voxel_size, 60)
    ref = ref_ext.aabb_intersect(rs, rd, pts[None].contiguous(), scene.voxel_size, 60)
    _cmp3(mine, ref, "degenerate rays vs reference CUDA")
    _cmp3(mine, _oracle_aabb(rs, rd, pts[None], scene.voxel_size, 60), "degenerate rays vs oracle")
    # KAT-3: the first ray crosses the 8 voxels of one x-row, entry/exit = (c_x -+ 0.125) + 3
    idx = mine[0][0, 0]
    hit = idx[idx >= 0]
    assert hit.numel() == 8
    cx = pts[hit.long(), 0]
    assert torch.equal(mine[1][0, 0, :8], (cx - 0.125) + 3.0)
    assert torch.equal(mine[2][0, 0, :8], (cx + 0.125) + 3.0)


def test_aabb_edge_shapes(cuda):
    scene = synthetic.make_scene("C2")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    e = torch.empty((1, 0, 3), device=cuda)
    out = ours.aabb_intersect(e, e, pts[None].contiguous(), 0.4, 60)
    assert out[0].shape == (1, 0, 60)
    o, d = helpers.rays_for("C2", 33, 5, cuda)   # ragged: not a multiple of anything
    far = pts + 100.0
    idx, dmin, dmax = ours.aabb_intersect(o[None].contiguous(), d[None].contiguous(), far[None].contiguous(), 0.4, 7)
    assert int((idx != -1).sum()) == 0 and float(dmin.abs().sum() + dmax.abs().sum()) == 0.0
    one = ours.aabb_intersect(o[None].contiguous(), d[None].contiguous(), pts[None, :1].contiguous(), 0.4, 3)
    assert one[0].shape == (1, 33, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ours.aabb_intersect(o[None].cpu(), d[None].cpu(), pts[None].cpu(), 0.4, 3)
    with pytest.raises(RuntimeError, match="must be a float tensor"):
        ours.aabb_intersect(o[None].double(), d[None].double(), pts[None].double(), 0.4, 3)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        ours.aabb_intersect(o[None].expand(2, -1, -1), d[None].expand(2, -1, -1), pts[None].expand(2, -1, -1), 0.4, 3)


def test_aabb_level2_wrapper_matches_reference_wrapper(cuda, ref_ext):
    """fairnr.clib.aabb_ray_intersect semantics (ray tiling / padding / voxel replication) end to end."""
    scene = synthetic.make_scene("C1")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    o, d = helpers.rays_for("C1", 5000, 7, cuda)     # 5000 is not a multiple of G=2048: wrap padding
    S = 1   # the reference wrapper's points.expand(S*G, ...) only works for a single shape
    rs, rd, P = o[None].contiguous(), d[None].contiguous(), pts[None].contiguous()
    vs, mh = torch.tensor(scene.voxel_size, device=cuda), torch.tensor(60.0, device=cuda)   # 0-dim tensor scalars
    mine = clib.aabb_ray_intersect(vs, mh, P, rs, rd)
    ref = wrappers.aabb_ray_intersect(ref_ext, float(vs), int(mh), P, rs, rd)
    _cmp3(mine, ref, "Level-2 aabb_ray_intersect")
    assert mine[0].shape == (S, 5000, 60) and not mine[0].requires_grad


def _svo_inputs(name_or_pts, vs, cuda):
    pts = torch.from_numpy(name_or_pts).to(cuda)
    centers, children = helpers.easy_octree(pts, vs, ours.build_octree)
    return pts, centers.contiguous(), children.contiguous()


@pytest.mark.parametrize("times,n_max,n_rays", [(0, 60, 4096), (1, 90, 4096), (2, 135, 2048), (1, 4, 2048)])
def test_svo_level1_vs_reference_and_oracle(cuda, ref_ext, times, n_max, n_rays):
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    p, vs = synthetic.split_points(pts0, 0.4, times)
    pts, centers, children = _svo_inputs(p, vs, cuda)
    rs, rd = synthetic.camera_rays(64, 64, 2, device=cuda)
    rs = rs.expand_as(rd)[:, :n_rays].contiguous()
    rd = rd[:, :n_rays].contiguous()
    cB = centers[None].expand(2, -1, -1).contiguous()
    chB = children[None].expand(2, -1, -1).contiguous()
    mine = ours.svo_intersect(rs, rd, cB, chB, vs, n_max)
    ref = ref_ext.svo_intersect(rs, rd, cB, chB, vs, n_max)
    _cmp3(mine, ref, "svo ours vs reference CUDA (T=%d)" % centers.shape[0])
    inv = helpers.ref_rcp(rd)
    orc = oracle.svo_intersect(rs.cpu().numpy(), rd.cpu().numpy(), cB.cpu().numpy(), chB.cpu().numpy(), vs, n_max,
                               inv.cpu().numpy())
    _cmp3(mine, orc, "svo ours vs CPU oracle")
    _cmp3(ours.svo_intersect(rs, rd, centers, children, vs, n_max, shared_tree=True), mine, "shared vs replicated")
    assert int((mine[0] >= 0).sum()) > n_rays
    if n_max >= 60:
        # property: leaf k == voxel k, and (untruncated) the svo hit SET equals aabb's on the same centres
        a = ours.aabb_intersect(rs, rd, pts, vs, n_max, shared_points=True)[0]
        assert torch.equal(a.sort(-1)[0], mine[0].sort(-1)[0])


def test_svo_level2_wrapper_matches_reference_wrapper(cuda, ref_ext):
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    pts, centers, children = _svo_inputs(pts0, 0.4, cuda)
    rs, rd = synthetic.camera_rays(50, 50, 1, device=cuda)
    rs = rs.expand_as(rd).contiguous()
    mine = clib.svo_ray_intersect(0.4, 60, centers[None], children[None], rs, rd)
    ref = wrappers.svo_ray_intersect(ref_ext, 0.4, 60, centers[None], children[None], rs, rd)
    _cmp3(mine, ref, "Level-2 svo_ray_intersect")


@pytest.mark.parametrize("times,n_max,n_rays", [(0, 60, 4096), (1, 90, 4096), (2, 135, 4096), (2, 20, 2048)])
def test_aabb_sorted_and_hit_mask_match_reference_postprocessing(cuda, ref_ext, times, n_max, n_rays):
    """The fused epilogue == SparseVoxelEncoder.ray_intersect's masked_fill + sort + gather + any (encoder.py:519-524)
    applied to the reference kernel's output; the any-hit kernel == its `hits`."""
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    p, vs = synthetic.split_points(pts0, 0.4, times)
    pts = torch.from_numpy(p).to(cuda)
    rs, rd = synthetic.camera_rays(64, 64, 1, device=cuda)
    rs = rs.expand_as(rd)[:, :n_rays].contiguous()
    rd = rd[:, :n_rays].contiguous()
    ref = ref_ext.aabb_intersect(rs, rd, pts[None].contiguous(), vs, n_max)
    r_idx, r_min, r_max, r_hits = wrappers.sort_hits(*ref)
    idx, dmin, dmax, hits = ours.aabb_intersect_sorted(rs, rd, pts[None].contiguous(), vs, n_max, 10000.0)
    assert torch.equal(hits, r_hits)
    assert torch.equal(dmin, r_min), "sorted min_depth differs"
    # torch.sort is not stable: rows with equal entry depths may permute in the reference; compare the rest exactly
    ties = (r_min[..., 1:] == r_min[..., :-1]) & (r_idx[..., 1:] != -1)
    clean = ~ties.any(-1)
    assert clean.float().mean() > 0.9
    assert torch.equal(idx[clean], r_idx[clean]) and torch.equal(dmax[clean], r_max[clean])
    assert torch.equal(idx.sort(-1)[0], r_idx.sort(-1)[0])
    assert torch.equal(ours.aabb_hit_mask(rs, rd, pts, vs, shared_points=True), r_hits)
    if n_max >= 135:
        assert int((idx >= 0).sum(-1).max()) > 32      # exercises the bitonic path


def test_encoder_ray_intersect_matches_reference_pipeline(cuda, ref_ext):
    """SparseVoxelEncoder.ray_intersect (mirror) == reference wrapper + reference post-processing, incl. the x shift."""
    from nsvf_b200.encoder import SparseVoxelEncoder
    scene = synthetic.make_scene("C1")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=60).to(cuda)
    st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=cuda))
    assert torch.equal(st["voxel_center_xyz"][0, :, 0], enc.points[:, 0] + enc.voxel_size / 10)
    rs, rd = synthetic.camera_rays(40, 40, 2, radius=3.0, device=cuda)
    rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
    ray_start, ray_dir, inter, hits = enc.ray_intersect(rs, rd, st)
    ref = wrappers.aabb_ray_intersect(ref_ext, float(enc.voxel_size), int(enc.max_hits), st["voxel_center_xyz"],
                                      ray_start, ray_dir)
    r_idx, r_min, r_max, r_hits = wrappers.sort_hits(*ref)
    assert torch.equal(hits, r_hits) and torch.equal(inter["min_depth"], r_min)
    assert torch.equal(inter["intersected_voxel_idx"].sort(-1)[0], r_idx.sort(-1)[0])
    _, _, h2 = enc.ray_hit_mask(rs, rd, st)
    assert torch.equal(h2, r_hits)


def test_encoder_octree_ray_intersect_matches_reference_pipeline(cuda, ref_ext):
    """use_octree=True: svo kernel + the in-place sort kernel == reference svo wrapper + torch post-processing."""
    from nsvf_b200.encoder import SparseVoxelEncoder
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    p, vs = synthetic.split_points(pts0, 0.4, 1)
    enc = SparseVoxelEncoder(p, vs, max_hits=90, use_octree=True).to(cuda)
    st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=cuda))
    rs, rd = synthetic.camera_rays(48, 48, 2, radius=4.5, device=cuda)
    rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
    ray_start, ray_dir, inter, hits = enc.ray_intersect(rs, rd, st)
    ref = wrappers.svo_ray_intersect(ref_ext, float(enc.voxel_size), int(enc.max_hits),
                                     st["voxel_octree_center_xyz"], st["voxel_octree_children_idx"], ray_start, ray_dir)
    r_idx, r_min, r_max, r_hits = wrappers.sort_hits(*ref)
    assert torch.equal(hits, r_hits) and torch.equal(inter["min_depth"], r_min)
    ties = (r_min[..., 1:] == r_min[..., :-1]) & (r_idx[..., 1:] != -1)
    clean = ~ties.any(-1)
    assert torch.equal(inter["intersected_voxel_idx"][clean], r_idx[clean])
    assert torch.equal(inter["max_depth"][clean], r_max[clean]) and int(hits.sum()) > 1000


def test_svo_non_enclosing_tree_falls_back_to_reference_boxes(cuda, ref_ext):
    """The tight-box shortcut is only valid when every node box encloses its children's boxes; a tree that violates
    this (here: internal centres displaced, size markers shrunk) must still reproduce the reference traversal."""
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    pts, centers, children = _svo_inputs(pts0, 0.4, cuda)
    n = pts.shape[0]
    bad_c, bad_ch = centers.clone(), children.clone()
    internal = torch.arange(n, centers.shape[0] - 1, device=cuda)
    bad_c[internal[::3]] += 0.35                       # displaced boxes: some leaves fall outside their ancestors
    bad_ch[internal[1::5], 8] = torch.clamp(bad_ch[internal[1::5], 8] // 4, min=2)   # shrunken boxes
    rs, rd = synthetic.camera_rays(48, 48, 1, device=cuda)
    rs = rs.expand_as(rd).contiguous()
    mine = ours.svo_intersect(rs, rd, bad_c[None].contiguous(), bad_ch[None].contiguous(), 0.4, 60)
    ref = ref_ext.svo_intersect(rs, rd, bad_c[None].contiguous(), bad_ch[None].contiguous(), 0.4, 60)
    _cmp3(mine, ref, "svo on a non-enclosing tree")
    good = ours.svo_intersect(rs, rd, centers[None].contiguous(), children[None].contiguous(), 0.4, 60)
    assert not torch.equal(good[0], mine[0])           # the malformed tree really changes the answer


def test_ball_and_triangle_intersect_vs_reference(cuda, ref_ext):
    """The two clib entry points outside the NSVF path, for API completeness: same outputs as the reference kernels."""
    scene = synthetic.make_scene("C1")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    o, d = helpers.rays_for("C1", 2048, 31, cuda)
    rs, rd = o.view(2, -1, 3).contiguous(), d.view(2, -1, 3).contiguous()
    P = pts[None].expand(2, -1, -1).contiguous()
    mine = ours.ball_intersect(rs, rd, P, 0.2, 16)
    ref = ref_ext.ball_intersect(rs, rd, P, 0.2, 16)
    assert torch.equal(mine[0], ref[0]) and int((mine[0] >= 0).sum()) > 2048
    torch.testing.assert_close(mine[1], ref[1], rtol=helpers.RTOL, atol=1e-6)
    torch.testing.assert_close(mine[2], ref[2], rtol=helpers.RTOL, atol=1e-6)
    # a triangulated box surface: 12 faces per voxel of a small grid
    g = torch.Generator().manual_seed(0)
    c = pts[torch.randperm(scene.n, generator=g)[:40]]
    off = torch.tensor([[a, b, e] for a in (-1., 1.) for b in (-1., 1.) for e in (-1., 1.)], device=cuda) * 0.125
    corners = (c[:, None] + off[None])                                    # [40, 8, 3]
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = [(q[0], q[1], q[2]) for q in quads] + [(q[0], q[2], q[3]) for q in quads]
    faces = torch.stack([corners[:, list(t)].reshape(-1, 9) for t in tris], 1).reshape(1, -1, 9)   # [1, 480, 9]
    F2 = faces.expand(2, -1, -1).contiguous()
    mine = ours.triangle_intersect(rs, rd, F2, 0.05, 0.01, 12)
    ref = ref_ext.triangle_intersect(rs, rd, F2, 0.05, 0.01, 12)
    assert torch.equal(mine[0], ref[0]) and int((mine[0] >= 0).sum()) > 500
    torch.testing.assert_close(mine[1], ref[1], rtol=helpers.RTOL, atol=1e-6)
    torch.testing.assert_close(mine[2], ref[2], rtol=helpers.RTOL, atol=1e-6)
    inds, depth, uv = clib.triangle_ray_intersect(0.05, 0.01, 12, corners.reshape(-1, 3),
                                                  torch.arange(480 * 3, device=cuda).view(-1, 3) * 0 +
                                                  torch.tensor([0, 1, 2], device=cuda), rs[:1], rd[:1])
    assert inds.shape == (1, rs.shape[1], 12) and depth.shape == (1, rs.shape[1], 12, 3)


def test_aabb_sorted_on_an_empty_voxel_set(cuda):
    """Everything pruned (n == 0): the sorted intersection returns all -1 / MAX_DEPTH and hits False, like the reference's
    masked_fill + sort of an all-miss result (this used to raise in the middle of a training step)."""
    rs = torch.randn(1, 100, 3, device=cuda)
    rd = torch.nn.functional.normalize(torch.randn(1, 100, 3, device=cuda), dim=-1)
    idx, dmin, dmax, hits = ours.aabb_intersect_sorted(rs, rd, torch.zeros(0, 3, device=cuda), 0.25, 7, 10000.0,
                                                       shared_points=True)
    assert idx.shape == (1, 100, 7) and bool((idx == -1).all()) and not bool(hits.any())
    assert bool((dmin == 10000.0).all()) and bool((dmax == 10000.0).all())


def _both_paths(fn):
    """Run fn() with the lattice walk (default) and with the hierarchy kernels only (NSVF_AABB_NO_GRID)."""
    import os
    walk = fn()
    os.environ["NSVF_AABB_NO_GRID"] = "1"
    try:
        tree = fn()
    finally:
        del os.environ["NSVF_AABB_NO_GRID"]
    return walk, tree


def _awkward_rays(pts, vs, n, cuda, seed=11):
    """Camera-like rays plus origins inside the set, on cell corners / faces, axis-parallel and zero components."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * 4.5
    d = (torch.rand(n, 3, generator=g) * 2 - 1) - o
    k = n // 8
    o[:k] = torch.rand(k, 3, generator=g) - 0.5                                   # inside the voxel set
    pick = torch.randint(0, pts.shape[0], (2 * k,), generator=g)
    o[k:3 * k] = pts.cpu()[pick] + vs * 0.5                                        # exactly on cell corners
    axis = torch.eye(3)[torch.randint(0, 3, (k,), generator=g)] * (torch.randint(0, 2, (k, 1), generator=g) * 2 - 1)
    d[k:2 * k] = axis                                                               # along lattice edges
    d[3 * k:4 * k, 1] = 0.0
    d[4 * k:4 * k + 8, 2] = -0.0
    d[4 * k + 8] = 0.0                                                              # no direction at all
    d[4 * k + 9, 0] = float("nan")
    o[4 * k + 10, 1] = float("inf")
    d[4 * k + 11] = torch.tensor([1e-30, 1e-30, 1e-30])
    o[4 * k + 12] = torch.tensor([3e6, 0.0, 0.0])                                   # too far for the walk: scans all voxels
    d = torch.where(d.norm(dim=-1, keepdim=True) > 0, d / d.norm(dim=-1, keepdim=True), d)
    return o.to(cuda)[None].contiguous(), d.to(cuda)[None].contiguous()


@pytest.mark.parametrize("times,n_max", [(0, 60), (1, 90), (2, 135), (1, 6)])
def test_lattice_walk_equals_hierarchy_and_reference(cuda, ref_ext, times, n_max):
    """voxel_grid.cu (cells along the ray, exact test on the occupied ones) against the hierarchy kernels and the
    reference scan: all three modes, awkward rays, rows that overflow n_max."""
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-1.2] * 3, [1.2] * 3, 0.4))
    p, vs = synthetic.split_points(pts0, 0.4, times)
    p[:, 0] += np.float32(vs / 10)
    pts = torch.from_numpy(p).to(cuda)
    rs, rd = _awkward_rays(pts, vs, 4096, cuda)
    walk, tree = _both_paths(lambda: ours.aabb_intersect(rs, rd, pts, vs, n_max, shared_points=True))
    _cmp3(walk, tree, "index order: walk vs hierarchy")
    _cmp3(walk, ref_ext.aabb_intersect(rs, rd, pts[None].contiguous(), vs, n_max), "index order: walk vs reference CUDA")
    walk, tree = _both_paths(lambda: ours.aabb_intersect_sorted(rs, rd, pts, vs, n_max, 10000.0, shared_points=True))
    _cmp3(walk[:3], tree[:3], "sorted: walk vs hierarchy")
    assert torch.equal(walk[3], tree[3])
    walk_m, tree_m = _both_paths(lambda: ours.aabb_hit_mask(rs, rd, pts, vs, shared_points=True))
    assert torch.equal(walk_m, tree_m) and torch.equal(walk_m, walk[3])
    if n_max == 6:
        assert int((walk[0] >= 0).sum(-1).max()) == 6      # overflowing rows exist
    # a voxel set prepared once serves any number of queries (nsvf_aabb_prepare / _intersect_prepared)
    index = ours.AabbIndex(pts, vs, shared_points=True)
    for sl in (slice(0, 4096), slice(1000, 3000)):
        a = ours.aabb_intersect_sorted(rs[:, sl].contiguous(), rd[:, sl].contiguous(), pts, vs, n_max, 10000.0, index=index)
        _cmp3(a[:3], [t[:, sl] for t in walk[:3]], "prepared index vs one-shot")
        assert torch.equal(a[3], walk[3][:, sl])
        assert torch.equal(ours.aabb_hit_mask(rs[:, sl].contiguous(), rd[:, sl].contiguous(), pts, vs, index=index), walk_m[:, sl])
    # per-batch voxel sets (one lattice per set)
    B = 2
    ptsB = torch.stack([pts, pts + 0.37])
    rsB, rdB = rs.view(B, -1, 3).contiguous(), rd.view(B, -1, 3).contiguous()
    walk, tree = _both_paths(lambda: ours.aabb_intersect_sorted(rsB, rdB, ptsB, vs, n_max, 10000.0))
    _cmp3(walk[:3], tree[:3], "sorted, two voxel sets: walk vs hierarchy")


def test_voxel_sets_off_the_lattice_use_the_hierarchy(cuda, ref_ext):
    """Jittered centres, duplicate centres and a set too sparse for the cell budget are not lattices: the device-side
    check hands them to the hierarchy kernels, results stay the reference's."""
    scene = synthetic.make_scene("C2")
    pts, _, _ = helpers.scene_tensors(scene, cuda)
    o, d = helpers.rays_for("C2", 4096, 9, cuda)
    rs, rd = o[None].contiguous(), d[None].contiguous()
    g = torch.Generator().manual_seed(3)
    jitter = pts + (torch.rand(pts.shape, generator=g).to(cuda) - 0.5) * 0.1
    twice = torch.cat([pts, pts[:17]])
    sparse = torch.cat([pts, pts[:4] + 400.0 * scene.voxel_size])
    for name, p in (("jittered", jitter), ("duplicates", twice), ("sparse", sparse)):
        p = p[None].contiguous()
        _cmp3(ours.aabb_intersect(rs, rd, p, scene.voxel_size, 60), ref_ext.aabb_intersect(rs, rd, p, scene.voxel_size, 60),
              name + " vs reference CUDA")
        r_idx, r_min, r_max, r_hits = wrappers.sort_hits(*ref_ext.aabb_intersect(rs, rd, p, scene.voxel_size, 60))
        idx, dmin, dmax, hits = ours.aabb_intersect_sorted(rs, rd, p, scene.voxel_size, 60, 10000.0)
        assert torch.equal(hits, r_hits) and torch.equal(dmin, r_min) and torch.equal(idx.sort(-1)[0], r_idx.sort(-1)[0])
        assert torch.equal(ours.aabb_hit_mask(rs, rd, p, scene.voxel_size), r_hits)


def _svo_then_sort(rs, rd, centers, children, vs, n_max, **kw):
    idx, dmin, dmax = ours.svo_intersect(rs, rd, centers, children, vs, n_max, **kw)
    idx, dmin, dmax = idx.clone(), dmin.clone(), dmax.clone()
    hits = ours.sort_hits_by_depth(idx, dmin, dmax, 10000.0)
    return idx, dmin, dmax, hits


@pytest.mark.parametrize("times,n_max", [(0, 60), (1, 90), (2, 135), (1, 5)])
def test_svo_sorted_equals_traversal_then_sort(cuda, ref_ext, times, n_max):
    """nsvf_svo_intersect_sorted (rays walk the lattice of the leaves; DFS ranks order ties and decide truncation) ==
    the reference DFS traversal followed by the stable depth sort, entry for entry — awkward rays, overflowing rows,
    shared and per-batch trees — and == the same call with the lattice switched off."""
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-1.2] * 3, [1.2] * 3, 0.4))
    p, vs = synthetic.split_points(pts0, 0.4, times)
    pts, centers, children = _svo_inputs(p, vs, cuda)
    rs, rd = _awkward_rays(pts, vs, 4096, cuda)
    want = _svo_then_sort(rs, rd, centers, children, vs, n_max, shared_tree=True)
    walk, tree = _both_paths(lambda: ours.svo_intersect_sorted(rs, rd, centers, children, vs, n_max, 10000.0, shared_tree=True))
    _cmp3(walk[:3], want[:3], "svo sorted (walk) vs traversal + sort")
    _cmp3(tree[:3], want[:3], "svo sorted (lattice off) vs traversal + sort")
    assert torch.equal(walk[3], want[3]) and torch.equal(tree[3], want[3])
    assert int((walk[0] >= 0).sum()) > 4096
    if n_max == 5:
        assert int((walk[0] >= 0).sum(-1).max()) == 5
    # an octree prepared once (nsvf_svo_prepare) answers any number of ray batches
    index = ours.SvoIndex(centers, children, vs, shared_tree=True)
    for sl in (slice(0, 4096), slice(500, 2500)):
        a = ours.svo_intersect_sorted(rs[:, sl].contiguous(), rd[:, sl].contiguous(), centers, children, vs, n_max, 10000.0,
                                      index=index)
        _cmp3(a[:3], [t[:, sl] for t in want[:3]], "prepared octree vs one-shot")
        assert torch.equal(a[3], want[3][:, sl])
    # reference traversal + torch post-processing (tie rows as sets: torch.sort is not stable)
    ref = ref_ext.svo_intersect(rs, rd, centers[None].contiguous(), children[None].contiguous(), vs, n_max)
    r_idx, r_min, r_max, r_hits = wrappers.sort_hits(*ref)
    assert torch.equal(walk[3], r_hits) and torch.equal(walk[1], r_min)
    assert torch.equal(walk[0].sort(-1)[0], r_idx.sort(-1)[0])
    # one tree per batch row
    cB = torch.stack([centers, centers + 0.21]).contiguous()
    chB = children[None].expand(2, -1, -1).contiguous()
    rsB, rdB = rs.view(2, -1, 3).contiguous(), rd.view(2, -1, 3).contiguous()
    _cmp3(ours.svo_intersect_sorted(rsB, rdB, cB, chB, vs, n_max, 10000.0)[:3], _svo_then_sort(rsB, rdB, cB, chB, vs, n_max)[:3],
          "svo sorted, two trees")


def test_svo_sorted_on_trees_the_walk_must_refuse(cuda, ref_ext):
    """A tree whose boxes do not enclose their children, and a tree with a leaf that is not connected to the root: the
    first goes to the traversal kernel as a whole, the second keeps the walk but must not report the orphan."""
    pts0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    pts, centers, children = _svo_inputs(pts0, 0.4, cuda)
    n = pts.shape[0]
    rs, rd = synthetic.camera_rays(48, 48, 1, device=cuda)
    rs = rs.expand_as(rd).contiguous()
    bad_c, bad_ch = centers.clone(), children.clone()
    internal = torch.arange(n, centers.shape[0] - 1, device=cuda)
    bad_c[internal[::3]] += 0.35
    bad_ch[internal[1::5], 8] = torch.clamp(bad_ch[internal[1::5], 8] // 4, min=2)
    _cmp3(ours.svo_intersect_sorted(rs, rd, bad_c, bad_ch, 0.4, 60, 10000.0, shared_tree=True)[:3],
          _svo_then_sort(rs, rd, bad_c, bad_ch, 0.4, 60, shared_tree=True)[:3], "non-enclosing tree")
    # cut one leaf out of its parent's child list: the node stays in the arrays but the DFS can no longer reach it
    orphan_ch = children.clone()
    hit_leaf = int(ours.svo_intersect(rs, rd, centers, children, 0.4, 60, shared_tree=True)[0].max())
    rows, cols = (orphan_ch[:, :8] == hit_leaf).nonzero(as_tuple=True)
    assert rows.numel() == 1
    orphan_ch[rows[0], cols[0]] = -1
    got = ours.svo_intersect_sorted(rs, rd, centers, orphan_ch, 0.4, 60, 10000.0, shared_tree=True)
    _cmp3(got[:3], _svo_then_sort(rs, rd, centers, orphan_ch, 0.4, 60, shared_tree=True)[:3], "tree with an orphan leaf")
    assert int((got[0] == hit_leaf).sum()) == 0
