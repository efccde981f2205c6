"""GPU parity at the sizes of BASELINE.json configs[3] (C4: octree over ~0.6 M voxels + pruning with a sharded
keep mask) and configs[4] (C5: hierarchical inverse-CDF pass whose bins are the coarse samples), plus the full C3
pipeline against a plain-torch statement of the reference stages.  Size-independent properties where the reference
kernels would take too long."""
import numpy as np
import pytest
import torch

from oracle import wrappers
from nsvf_b200 import synthetic, clib, geometry, ops
from nsvf_b200.clib import _ext as ours
from nsvf_b200.encoder import SparseVoxelEncoder
from tests import helpers

pytestmark = pytest.mark.gpu


def test_c4_octree_intersection_vs_reference(cuda, ref_ext):
    scene = synthetic.make_scene("C4")                       # 13^3 shell split 3x: ~0.59 M voxels, voxel 0.05
    assert scene.n > 500000
    pts = torch.from_numpy(scene.points).to(cuda)
    centers, children = geometry.build_easy_octree(pts, scene.voxel_size / 2.0)   # host builder, ~1.3 M nodes
    assert (children[:, 8] == 1).sum() == scene.n and children.shape[0] > 600000
    rs, rd = synthetic.camera_rays(120, 120, 1, radius=4.5, seed=3, device=cuda)
    rs = rs.expand_as(rd).contiguous()
    P = scene.max_hits                                        # 202
    mine = clib.svo_ray_intersect(scene.voxel_size, P, centers[None], children[None], rs, rd)
    ref = ref_ext.svo_intersect(rs, rd, centers[None].contiguous(), children[None].contiguous(), scene.voxel_size, P)
    for a, b, nm in zip(mine, ref, ("idx", "min_depth", "max_depth")):
        assert torch.equal(a, b), "svo %s differs from the reference kernel at C4 scale" % nm
    n_hits = (mine[0] >= 0).sum(-1)
    assert int(n_hits.max()) > 60
    # property: the octree path and the index-ordered path find the same voxel SET (both untruncated here)
    # (on the SAME centres: leaf k of the octree is voxel k, its centre is coords * half_voxel + residual)
    leaf_pts = centers[:scene.n].contiguous()
    a = ours.aabb_intersect(rs, rd, leaf_pts, scene.voxel_size, P, shared_points=True)[0]
    assert torch.equal(a.sort(-1)[0], mine[0].sort(-1)[0])
    # property: sorted-by-depth output is non-decreasing and its any() equals the any-hit kernel
    idx, dmin, dmax, hits = ours.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, P, 1e4, shared_points=True)
    assert bool((dmin[..., 1:] >= dmin[..., :-1]).all()) and torch.equal(hits, (idx >= 0).any(-1))
    assert torch.equal(ours.aabb_hit_mask(rs, rd, pts, scene.voxel_size, shared_points=True), hits)


def test_c4_pruning_sharded_mask_equals_single_rank(cuda):
    scene = synthetic.make_scene("C3")
    enc = SparseVoxelEncoder(scene.points[:4096], scene.voxel_size, max_hits=135).to(cuda)
    field = lambda inp, outputs: {"sigma": inp["emb"][:, 0] * 9 + inp["emb"][:, 3] * 4 - 0.3}
    keep_full, score_full = enc._prune_scores(field, 0.5, bits=8)
    from nsvf_b200 import dist as nd
    parts = []
    for rank in range(3):                                     # what each of 3 ranks would compute
        lo, hi = nd.shard_range(4096, rank, 3)
        k, s = enc._prune_scores(field, 0.5, bits=8, lo=lo, hi=hi)
        assert k.numel() == hi - lo
        parts.append(k)
    assert torch.equal(torch.cat(parts), keep_full) and 0 < int(keep_full.sum()) < 4096
    ref_scores = enc.get_scores(field, bits=8)                # [n, 512] like the reference's get_scores(bits=8)
    assert ref_scores.shape == (4096, 512)
    helpers.assert_close_scaled(score_full, ref_scores.min(-1)[0], what="fused min-score vs get_scores")


def test_c5_hierarchical_sampling_vs_reference(cuda, ref_ext):
    """Fine pass of nerf.py:64-79: bins = coarse samples (depth -+ dist/2), probs = (w + 1e-5)/sum, fixed number of
    fine samples; max_hits = coarse max_len (hundreds of bins per ray)."""
    scene = synthetic.make_scene("C5")
    pts = torch.from_numpy(scene.points).to(cuda)
    rs, rd = synthetic.camera_rays(54, 96, 1, radius=9.0, seed=1, device=cuda)     # 1920x1080 / 20
    rs = rs.expand_as(rd).contiguous()
    idx, dmin, dmax, hits = ours.aabb_intersect_sorted(rs, rd, pts, scene.voxel_size, scene.max_hits, 1e4,
                                                       shared_points=True)
    sel = hits[0]
    idx, dmin, dmax = idx[0][sel].contiguous(), dmin[0][sel].contiguous(), dmax[0][sel].contiguous()
    assert idx.shape[0] > 500
    probs, steps = wrappers.probs_and_steps(idx, dmin, dmax, scene.step_size)
    c_idx, c_depth, c_dist = wrappers.mask_samples(*clib.inverse_cdf_sampling(idx, dmin, dmax, probs, steps, -1, True))
    assert c_idx.shape[1] > 150                                # long rays: hundreds of coarse samples
    w = torch.rand_like(c_depth) * c_idx.ne(-1)
    fine = {"min_depth": (c_depth - c_dist * .5).contiguous(), "max_depth": (c_depth + c_dist * .5).contiguous(),
            "idx": c_idx.contiguous()}
    safe = w + 1e-5
    f_probs = (safe / safe.sum(-1, keepdim=True)).contiguous()
    f_steps = torch.full((c_idx.shape[0],), 64.0, device=cuda)
    torch.manual_seed(5)
    mine = clib.inverse_cdf_sampling(fine["idx"], fine["min_depth"], fine["max_depth"], f_probs, f_steps, -1, False)
    torch.manual_seed(5)
    ref = wrappers.inverse_cdf_sampling(ref_ext, fine["idx"], fine["min_depth"], fine["max_depth"], f_probs, f_steps,
                                        -1, False)
    for a, b, nm in zip(mine, ref, ("sampled_idx", "sampled_depth", "sampled_dists")):
        assert a.shape == b.shape and torch.equal(a, b), "hierarchical %s differs from the reference" % nm
    assert int(mine[0].ne(-1).sum(-1).max()) >= 64


def test_c3_pipeline_matches_torch_statement_of_reference(cuda):
    """Full eval pipeline on a C3-shaped scene (intersect -> sample -> interpolate -> field -> composite with early
    termination off) against the same stages written in plain torch (oracle/wrappers.py) on identical samples."""
    scene = synthetic.make_scene("C3")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits).to(cuda).eval()
    from nsvf_b200.renderer import VolumeRenderer
    from nsvf_b200.field import TrivialField
    from nsvf_b200.pipeline import NSVFPipeline
    pipe = NSVFPipeline(enc, TrivialField(), VolumeRenderer(chunk_size=64), pixel_per_view=0).to(cuda).eval()
    rs, rd = synthetic.camera_rays(96, 96, 1, radius=4.5, seed=2, device=cuda)
    with torch.no_grad():
        out = pipe(rs[None, :, None, 0, :].contiguous(), rd[None].contiguous())
        s = out["samples"]
        sidx, sdep, sdist = s["sampled_point_voxel_idx"], s["sampled_point_depth"], s["sampled_point_distance"]
        st = enc.precompute(id=None)
        hit = out["hits"]
        o = rs.expand_as(rd).reshape(-1, 3)[hit]
        d = rd.reshape(-1, 3)[hit]
        # the pipeline hands trimmed rows to the renderer: slots beyond sampled_point_count are never written
        mask = torch.arange(sidx.shape[1], device=cuda)[None] < s["sampled_point_count"][:, None]
        sdep, sdist = sdep.masked_fill(~mask, 10000.0), sdist.masked_fill(~mask, 0.0)
        pipe.padded_samples = True                       # the reference's padded [N, max_len] rows: same results
        out_p = pipe(rs[None, :, None, 0, :].contiguous(), rd[None].contiguous())
        K = out_p["samples"]["sampled_point_voxel_idx"].shape[1]
        assert torch.equal(out_p["samples"]["sampled_point_voxel_idx"].ne(-1), mask[:, :K]) and not bool(mask[:, K:].any())
        assert torch.equal(out_p["samples"]["sampled_point_depth"], sdep[:, :K])
        for key in ("colors", "missed", "depths"):
            assert torch.equal(out_p[key], out[key]), key
        xyz = (o[:, None] + d[:, None] * sdep[..., None])[mask]
        emb = wrappers.trilinear_torch(sidx[mask].long(), xyz, st["voxel_vertex_idx"], st["voxel_center_xyz"],
                                       st["voxel_vertex_emb"], enc.voxel_size)
        sigma, tex_c = emb[:, 0] * 4 + 1, torch.tanh(emb[:, 1:4])
        fe = torch.zeros_like(sdep).masked_scatter(mask, torch.relu(sigma) * sdist[mask] * 7.0)
        tex = torch.zeros(*sdep.shape, 3, device=cuda).masked_scatter(mask[..., None].expand(-1, -1, 3), tex_c)
        probs, depth, missed, colors = wrappers.composite_torch(fe, tex, sdep)
    n = int(hit.sum())
    assert n > 3000 and out["ae"] == int(mask.sum())
    helpers.assert_close_scaled(out["colors"][hit] - out["missed"][hit][:, None], colors, what="pipeline colours")
    helpers.assert_close_scaled(1 - out["missed"][hit], 1 - missed, what="pipeline opacity")
    assert bool((out["missed"][~hit] == 1).all())


def test_c5_hierarchical_pipeline_runs_and_is_consistent(cuda):
    """Two-pass (coarse + importance-sampled fine) pipeline, training mode with gradients into the embeddings."""
    from nsvf_b200.renderer import VolumeRenderer
    from nsvf_b200.field import TrivialField
    from nsvf_b200.pipeline import NSVFPipeline
    scene = synthetic.make_scene("C1")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=60).to(cuda)
    pipe = NSVFPipeline(enc, TrivialField(), VolumeRenderer(chunk_size=64), pixel_per_view=256,
                        hierarchical_sampling=True, fixed_fine_num_samples=48).to(cuda).train()
    rs, rd = synthetic.camera_rays(64, 64, 2, radius=3.0, seed=4, device=cuda)
    torch.manual_seed(0)
    out = pipe(rs[None, :, None, 0, :].contiguous(), rd[None].contiguous())
    assert out["colors"].shape == (512, 3) and "coarse" in out and out["ae"] > 0
    s = out["samples"]["sampled_point_voxel_idx"]
    assert int(s.ne(-1).sum(-1).min()) >= 48                 # every marched ray got its fine samples
    (out["colors"].sum() + out["missed"].sum()).backward()
    g = enc.values.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    # the coarse pass of the two-pass pipeline is the single-pass pipeline (same seed: same pixel draw, same sampling and
    # sigma noise), nerf.py:64-79 only adds the second pass
    single = NSVFPipeline(enc, TrivialField(), VolumeRenderer(chunk_size=64), pixel_per_view=256).to(cuda).train()
    single.padded_samples = True
    torch.manual_seed(0)
    out1 = single(rs[None, :, None, 0, :].contiguous(), rd[None].contiguous())
    assert torch.equal(out1["sampled"], out["sampled"])
    hit = out["hits"]
    helpers.assert_close_scaled(out1["missed"][hit], out["coarse"]["missed"], what="coarse pass == single-pass pipeline")
    assert bool(((out["missed"] >= -1e-5) & (out["missed"] <= 1 + 1e-5)).all())
