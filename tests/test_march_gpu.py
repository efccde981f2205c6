"""GPU tests of the device-side ray-marching plan (csrc/march.cu, renderer._forward_chunk_plan): the chunk schedule,
trimmed-row compaction / epilogue / compositing and their backward, against

  * the general path of the same renderer (count / scan / fill compaction + index_put + dense compositing), which the
    golden tests pin to the unmodified reference renderer (tests/test_golden_gpu.py), and
  * a plain-torch statement of the reference's forward_chunk loop (fairnr/modules/renderer.py:135-232).
"""
import math
import os

import numpy as np
import pytest
import torch

from nsvf_b200 import _lib, clib, synthetic
from nsvf_b200.encoder import SparseVoxelEncoder
from nsvf_b200.renderer import VolumeRenderer
from oracle import wrappers
from tests import helpers

pytestmark = pytest.mark.gpu


def _field(inputs, outputs=("sigma", "texture")):
    emb = inputs["emb"]
    if "sigma" in outputs:
        inputs["sigma"] = emb[:, 0] * 6 + emb[:, 5] * 3 + 0.5
    if "texture" in outputs:
        inputs["texture"] = torch.tanh(emb[:, 1:4] * 2)
    return inputs


def _scene_samples(cuda, n_rays=3000, seed=3, trimmed=False, train=False):
    scene = synthetic.make_scene("C1")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits).to(cuda)
    enc.train(train)
    o, d = synthetic.random_rays(n_rays, seed=seed)
    rs, rd = torch.from_numpy(o).to(cuda)[None, :, None], torch.from_numpy(d).to(cuda)[None, :, None]
    st = enc.precompute(id=None)
    rs_f, rd_f, inter, hits = enc.ray_intersect(rs.reshape(1, -1, 1, 3), rd.reshape(1, -1, 1, 3),
                                                {k: v[None] for k, v in st.items()})
    hits = hits.reshape(-1)
    inter = {k: v.reshape(-1, v.size(-1))[hits] for k, v in inter.items()}
    dists = (inter["max_depth"] - inter["min_depth"]).masked_fill(inter["intersected_voxel_idx"].eq(-1), 0)
    inter["probs"] = dists / dists.sum(-1, keepdim=True)
    inter["steps"] = dists.sum(-1) / enc.step_size
    torch.manual_seed(11)
    samples = enc.ray_sample(inter, trimmed=trimmed)
    return enc, st, rs_f.reshape(-1, 3)[hits], rd_f.reshape(-1, 3)[hits], samples, inter


def _render(ren, enc, st, rs, rd, samples, general, seed=5):
    if general:
        os.environ["NSVF_RENDER_GENERAL"] = "1"
    else:
        os.environ.pop("NSVF_RENDER_GENERAL", None)
    try:
        torch.manual_seed(seed)
        enc.values.weight.grad = None
        res = ren(enc, _field, rs, rd, samples, st)
        loss = (res["colors"] ** 2).sum() + res["missed"].sum() * 0.3 + res["depths"].sum() * 0.1 \
            + (res["probs"] * torch.linspace(0, 1, res["probs"].shape[1], device=rs.device)).sum() * 0.05
        if loss.requires_grad:
            loss.backward()
        return res, None if enc.values.weight.grad is None else enc.values.weight.grad.clone()
    finally:
        os.environ.pop("NSVF_RENDER_GENERAL", None)


@pytest.mark.parametrize("train,tol,chunk", [(False, 0.0, 64), (False, 0.0, 2), (False, 0.05, 2), (False, 0.3, 1),
                                             (True, 0.0, 8), (True, 0.1, 3)])
def test_plan_path_equals_general_path(cuda, train, tol, chunk):
    enc, st, rs, rd, samples, _ = _scene_samples(cuda, train=train)
    ren = VolumeRenderer(chunk_size=chunk, valid_chunk_size=chunk, discrete_regularization=train,
                         raymarching_tolerance=tol).train(train)
    a, ga = _render(ren, enc, st, rs, rd, samples, general=False)
    b, gb = _render(ren, enc, st, rs, rd, samples, general=True)
    assert a["ae"] == b["ae"], "the two paths sent different sets of samples to the field"
    for name in ("probs", "depths", "colors", "max_depths", "min_depths", "missed"):
        helpers.assert_close_scaled(a[name], b[name], what="%s (plan vs general)" % name)
    assert (ga is None) == (gb is None)
    if ga is not None:
        helpers.assert_close_scaled(ga, gb, what="d loss / d values (plan vs general)")


def test_trimmed_rows_equal_padded_rows(cuda):
    """The trimmed sampler output (no padding written, no max_len sync) feeds the plan path to the same results."""
    enc, st, rs, rd, padded, inter = _scene_samples(cuda, trimmed=False)
    torch.manual_seed(11)
    trimmed = enc.ray_sample(inter, trimmed=True)
    n = trimmed["sampled_point_count"].long()
    K = padded["sampled_point_voxel_idx"].shape[1]
    assert int(n.max()) == K, "max_len of the padded rows must be the longest trimmed row"
    mask = torch.arange(K, device=cuda)[None] < n[:, None]
    assert torch.equal(padded["sampled_point_voxel_idx"].ne(-1), mask)
    for key in ("sampled_point_voxel_idx", "sampled_point_depth", "sampled_point_distance"):
        assert torch.equal(padded[key][mask], trimmed[key][:, :K][mask]), key
    assert float(padded["sampled_point_depth"][~mask].min()) == 10000.0 and float(padded["sampled_point_distance"].min()) >= 0
    ren = VolumeRenderer(chunk_size=4, valid_chunk_size=4, raymarching_tolerance=0.05).eval()
    a, _ = _render(ren, enc, st, rs, rd, padded, general=False)
    b, _ = _render(ren, enc, st, rs, rd, trimmed, general=False)
    assert a["ae"] == b["ae"]
    for name in ("depths", "colors", "max_depths", "min_depths", "missed"):
        assert torch.equal(a[name], b[name]), name
    assert torch.equal(a["probs"], b["probs"][:, :K]) and float(b["probs"][:, K:].abs().max()) == 0


def test_plan_schedule_matches_reference_rule(cuda):
    """The window list the device publishes equals the reference's flush rule (renderer.py:157-158) replayed on the
    host from the per-column counts."""
    L, p = _lib.load(), _lib.ptr
    g = torch.Generator().manual_seed(0)
    for B, K, chunk in ((1000, 37, 1500), (5000, 130, 4096), (300, 5, 100000), (64, 300, 7)):
        lens = torch.randint(0, K + 1, (B,), generator=g).int().to(cuda)
        plan = torch.zeros(L.nsvf_march_plan_bytes(B, K) // 4 + 2, dtype=torch.int32, device=cuda)
        info = torch.zeros(16 + 3 * (K + 1), dtype=torch.int32).pin_memory()
        _lib.check(L.nsvf_march_begin(_lib.current_stream(cuda), B, K, chunk, p(lens), None, 1, p(plan),
                                      info.data_ptr(), info.numel()))
        torch.cuda.synchronize()
        got = info[16: 16 + 3 * int(info[5])].view(-1, 3).tolist()
        counts = (torch.arange(K)[None] < lens.cpu()[:, None]).sum(0).tolist()
        want, size, start = [], 0, 0
        for i in range(K + 1):
            if (i == K or size + counts[i] > chunk) and i > start:
                if size > 0:
                    want.append([start, i, size])
                start, size = i, 0
            if i < K:
                size += counts[i]
        assert got == want, (B, K, chunk)
        assert int(info[8]) == int(lens.sum())


def test_rows_with_holes_fall_back_to_the_general_path(cuda):
    enc, st, rs, rd, samples, _ = _scene_samples(cuda, n_rays=500)
    idx = samples["sampled_point_voxel_idx"].contiguous().clone()
    dep, dst = samples["sampled_point_depth"].contiguous().clone(), samples["sampled_point_distance"].contiguous().clone()
    rows = idx[:, 2].ne(-1).nonzero()[:40, 0]
    idx[rows, 1], dep[rows, 1], dst[rows, 1] = -1, 10000.0, 0.0          # a hole in the middle of the row
    holed = {"sampled_point_voxel_idx": idx, "sampled_point_depth": dep, "sampled_point_distance": dst}
    ren = VolumeRenderer(chunk_size=1, valid_chunk_size=1, raymarching_tolerance=0.0).eval()
    a, _ = _render(ren, enc, st, rs, rd, holed, general=False)
    b, _ = _render(ren, enc, st, rs, rd, holed, general=True)
    assert a["ae"] == b["ae"] == int(idx.ne(-1).sum())
    for name in ("probs", "depths", "colors", "missed"):
        assert torch.equal(a[name], b[name]), name


def test_plan_matches_torch_statement_of_reference_loop(cuda):
    """forward_chunk against the reference loop written in plain torch (renderer.py:135-232) on the same inputs,
    with early termination: identical evaluation count, results within the fp32 tolerance."""
    enc, st, rs, rd, samples, _ = _scene_samples(cuda, n_rays=2000, seed=9)
    tol, chunk = 0.2, 2
    ren = VolumeRenderer(chunk_size=chunk, valid_chunk_size=chunk, raymarching_tolerance=tol).eval()
    with torch.no_grad():
        res = ren(enc, _field, rs, rd, samples, st)
        sidx, depth, dists = (samples["sampled_point_voxel_idx"].long(), samples["sampled_point_depth"],
                              samples["sampled_point_distance"])
        B, K = sidx.shape
        hits = sidx.ne(-1).long()
        fe_full, tex_full = torch.zeros(B, K, device=cuda), torch.zeros(B, K, 3, device=cuda)
        acc, early, evals, size, start = torch.zeros(B, device=cuda), None, 0, 0, 0
        for i in range(K + 1):
            if (i == K or size + int(hits[:, i].sum()) > chunk * 1024) and i > start:
                mask = sidx[:, start:i].ne(-1)
                if early is not None:
                    mask = mask & ~early[:, None]
                if int(mask.sum()) > 0:
                    xyz = rs[:, None] + rd[:, None] * depth[:, start:i, None]
                    inp = enc({"sampled_point_voxel_idx": sidx[:, start:i][mask], "sampled_point_xyz": xyz[mask],
                               "sampled_point_ray_direction": rd[:, None].expand(B, i - start, 3)[mask],
                               "sampled_point_distance": dists[:, start:i][mask]}, st)
                    out = _field(inp)
                    fe = torch.relu(out["sigma"]) * inp["dists"] * 7.0
                    fe_full[:, start:i] = torch.zeros(B, i - start, device=cuda).masked_scatter(mask, fe)
                    tex_full[:, start:i] = torch.zeros(B, i - start, 3, device=cuda).masked_scatter(
                        mask[..., None].expand(-1, -1, 3), out["texture"])
                    acc += fe_full[:, start:i].sum(1)
                    early = acc > -math.log(tol)
                    hits[early] *= 0
                    evals += int(mask.sum())
                start, size = i, 0
            if i < K:
                size += int(hits[:, i].sum())
        probs, dep, missed, colors = wrappers.composite_torch(fe_full, tex_full, depth)
    assert res["ae"] == evals
    helpers.assert_close_scaled(res["probs"], probs, what="probs")
    helpers.assert_close_scaled(res["colors"], colors, what="colors")
    helpers.assert_close_scaled(res["depths"], dep, what="depths")
    helpers.assert_close_scaled(res["max_depths"], depth.masked_fill(hits.eq(0), -1).max(1).values, what="max_depths")
    helpers.assert_close_scaled(res["min_depths"], depth.min(1).values, what="min_depths")


@pytest.mark.parametrize("deterministic", [True, False])
def test_lazy_sampling_blocks_equal_eager_samples(cuda, deterministic):
    """nsvf_inverse_cdf_plan / _block (on-demand sampling) reproduce the eager sampler bit for bit: same per-ray counts,
    and every block of the slot-major planes holds exactly the eager samples (all rays alive, all columns)."""
    L, p = _lib.load(), _lib.ptr
    enc, st, rs, rd, _, inter = _scene_samples(cuda, n_rays=4000, seed=13, train=not deterministic)
    args = (inter["intersected_voxel_idx"], inter["min_depth"], inter["max_depth"], inter["probs"], inter["steps"])
    torch.manual_seed(21)
    eidx, edep, edst, elen, _ = clib.inverse_cdf_sampling_rows(*args, -1, deterministic, trimmed=True)
    torch.manual_seed(21)
    lz = clib.inverse_cdf_sampling_lazy(*args, -1, deterministic)
    assert torch.equal(lz["sampled_point_count"], elen)
    meta = lz["lazy_meta"].tolist()
    assert meta[0] == int(elen.max()) and meta[2] == 0
    B, K = elen.numel(), lz["lazy_max_steps"]
    assert K == eidx.shape[1]
    ldb = L.nsvf_march_plane_stride(B)
    idxT = torch.full((K, ldb), -7, dtype=torch.int32, device=cuda)
    depT = torch.full((K, ldb), -7.0, device=cuda)
    dstT = torch.full((K, ldb), -7.0, device=cuda)
    noise = lz.get("lazy_noise", None)
    for k0 in range(0, K, 96):          # blocks that do not line up with the kernel's internal 64-column blocks
        _lib.check(L.nsvf_inverse_cdf_block(
            _lib.current_stream(cuda), B, inter["min_depth"].shape[1], K, -1.0, k0, min(K, k0 + 96), None,
            p(lz["sampled_point_count"]), p(lz["lazy_quirk"]), p(lz["lazy_pts_idx"]), p(lz["lazy_min_depth"]),
            p(lz["lazy_max_depth"]), p(noise), K, 0.5, p(lz["lazy_probs"]), p(lz["lazy_steps"]), 10000.0, p(idxT), p(depT),
            p(dstT)))
    mask = torch.arange(K, device=cuda)[None] < elen[:, None]
    assert torch.equal(idxT[:, :B].t()[mask], eidx[mask])
    assert torch.equal(depT[:, :B].t()[mask], edep[mask]) and torch.equal(dstT[:, :B].t()[mask], edst[mask])
    assert bool((idxT[:, :B].t()[~mask] == -7).all()), "nothing beyond a ray's samples may be written"
    # the resumable serial sampler (blocks in increasing order, state parked between calls); from the third block on
    # every fourth ray is flagged as stopped and must not be touched any more
    idxS, depS, dstS = torch.full_like(idxT, -7), torch.full_like(depT, -7.0), torch.full_like(dstT, -7.0)
    state = torch.empty(L.nsvf_inverse_cdf_stream_state_bytes(B), dtype=torch.uint8, device=cuda)
    stop = torch.zeros(B, dtype=torch.uint8, device=cuda)
    edges = [0, K // 7, K // 3, K // 3 + 1, (2 * K) // 3, K]
    for n, (k0, k1) in enumerate(zip(edges[:-1], edges[1:])):
        if n == 2:
            stop[::4] = 1
        _lib.check(L.nsvf_inverse_cdf_stream(
            _lib.current_stream(cuda), B, inter["min_depth"].shape[1], K, -1.0, k0, k1, p(stop) if n >= 2 else None,
            p(lz["sampled_point_count"]), p(lz["lazy_quirk"]), p(lz["lazy_pts_idx"]), p(lz["lazy_min_depth"]),
            p(lz["lazy_max_depth"]), p(noise), K, 0.5, p(lz["lazy_probs"]), p(lz["lazy_steps"]), 10000.0, p(state),
            p(idxS), p(depS), p(dstS)))
    want = mask & ~(stop.bool()[:, None] & (torch.arange(K, device=cuda)[None] >= edges[2]))
    assert K > 20 and int(want.sum()) < int(mask.sum())
    assert torch.equal(idxS[:, :B].t()[want], eidx[want])
    assert torch.equal(depS[:, :B].t()[want], edep[want]) and torch.equal(dstS[:, :B].t()[want], edst[want])
    assert bool((idxS[:, :B].t()[~want] == -7).all()), "stopped rays and positions beyond a ray's samples stay untouched"


def test_lazy_rendering_equals_eager_rendering(cuda):
    """Rendering with early termination on lazily sampled rays == on eagerly sampled (trimmed) rays, exactly."""
    enc, st, rs, rd, _, inter = _scene_samples(cuda, n_rays=5000, seed=17)
    eager = enc.ray_sample(inter, trimmed=True)
    lazy = enc.ray_sample(inter, lazy=True)
    for tol, chunk in ((0.05, 2), (0.3, 1), (0.0, 4)):
        ren = VolumeRenderer(chunk_size=chunk, valid_chunk_size=chunk, raymarching_tolerance=tol).eval()
        a, _ = _render(ren, enc, st, rs, rd, eager, general=False)
        b, _ = _render(ren, enc, st, rs, rd, lazy, general=False)
        assert a["ae"] == b["ae"]
        for name in ("depths", "colors", "max_depths", "min_depths", "missed", "probs"):
            assert torch.equal(a[name], b[name]), (name, tol)


@pytest.mark.parametrize("trimmed,block", [(False, "64"), (True, "64"), (True, "32")])
def test_compaction_queued_ahead_of_the_readback_changes_nothing(cuda, trimmed, block):
    """With early termination the next window's compaction is launched from the device-side schedule before the host
    has read it (nsvf_march_compact start = -1); switching that off, or making the plane blocks so small that the
    launch-ahead has to be repeated, gives bit-identical frames and evaluation counts."""
    enc, st, rs, rd, samples, inter = _scene_samples(cuda, n_rays=2500, seed=4, trimmed=trimmed)
    ren = VolumeRenderer(chunk_size=2, valid_chunk_size=2, raymarching_tolerance=0.1).eval()
    out = {}
    os.environ["NSVF_PLANE_BLOCK"] = block
    try:
        for mode in ("1", "0"):
            os.environ["NSVF_MARCH_AHEAD"] = mode
            with torch.no_grad():
                out[mode] = ren(enc, _field, rs, rd, samples, st)
    finally:
        os.environ.pop("NSVF_MARCH_AHEAD", None)
        os.environ.pop("NSVF_PLANE_BLOCK", None)
    assert out["1"]["ae"] == out["0"]["ae"] and out["1"]["ae"] > 0
    for name in ("probs", "depths", "colors", "max_depths", "min_depths", "missed"):
        assert torch.equal(out["1"][name], out["0"][name]), name


def test_encoder_window_fn_equals_forward(cuda):
    """SparseVoxelEncoder.window_fn (the lean per-window entry the renderer uses for inference) returns what forward()
    returns, and declines when autograd is on."""
    enc, st, rs, rd, samples, _ = _scene_samples(cuda, n_rays=500, seed=8)
    sidx = samples["sampled_point_voxel_idx"]
    mask = sidx.ne(-1)
    vox = sidx[mask].int().contiguous()
    depth = samples["sampled_point_depth"][mask]
    dirs = rd[:, None].expand(-1, sidx.shape[1], -1)[mask].contiguous()
    xyz = (rs[:, None].expand(-1, sidx.shape[1], -1)[mask] + dirs * depth[:, None]).contiguous()
    dists = samples["sampled_point_distance"][mask].contiguous()
    assert enc.window_fn(st, torch.cuda.current_stream(cuda).cuda_stream) is None      # grad mode: general forward
    with torch.no_grad():
        fn = enc.window_fn(st, torch.cuda.current_stream(cuda).cuda_stream)
        a = fn(vox, xyz, dirs, dists)
        b = enc({"sampled_point_voxel_idx": vox, "sampled_point_xyz": xyz, "sampled_point_ray_direction": dirs,
                 "sampled_point_distance": dists}, st)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
