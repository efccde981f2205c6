"""GPU parity of the fused torch-level stages: trilinear embedding interpolation (fwd + scatter-add bwd) and
alpha compositing (fwd + fused bwd) against (1) the CPU oracle and (2) a plain PyTorch fp32 statement of the
reference code evaluated on the same device, whose autograd is the reference's backward."""
import numpy as np
import pytest
import torch

import oracle
from oracle import wrappers
from nsvf_b200 import synthetic, ops
from tests import helpers

pytestmark = pytest.mark.gpu


def _samples(scene, M, seed, cuda, runs=True):
    rng = np.random.RandomState(seed)
    if runs:   # ray-marched order: runs of consecutive samples in the same voxel
        vox = np.repeat(rng.randint(0, scene.n, size=M // 6 + 1), rng.randint(1, 12, size=M // 6 + 1))[:M]
        if len(vox) < M:
            vox = np.concatenate([vox, rng.randint(0, scene.n, size=M - len(vox))])
    else:
        vox = rng.randint(0, scene.n, size=M)
    xyz = scene.points[vox] + rng.uniform(-0.5, 0.5, size=(M, 3)).astype("float32") * scene.voxel_size
    return torch.from_numpy(vox.astype("int64")).to(cuda), torch.from_numpy(xyz.astype("float32")).to(cuda)


@pytest.mark.parametrize("name,M,D", [("C1", 50000, 32), ("C2", 20001, 32), ("C1", 5000, 16), ("C1", 777, 48)])
def test_trilinear_forward_backward(cuda, name, M, D):
    scene = synthetic.Scene(synthetic.make_scene(name).points, synthetic.make_scene(name).voxel_size, embed_dim=D)
    pts, feats, values = helpers.scene_tensors(scene, cuda)
    vox, xyz = _samples(scene, M, 0, cuda)
    values = values.clone().requires_grad_(True)
    xyz = xyz.clone().requires_grad_(True)
    emb = ops.trilinear_embed(vox, xyz, feats, pts, values, scene.voxel_size)
    v2 = values.detach().clone().requires_grad_(True)
    x2 = xyz.detach().clone().requires_grad_(True)
    ref = wrappers.trilinear_torch(vox, x2, feats, pts, v2, scene.voxel_size)
    torch.testing.assert_close(emb, ref, rtol=helpers.RTOL, atol=1e-6)
    orc = oracle.trilinear_fwd(vox.cpu().numpy(), xyz.detach().cpu().numpy(), scene.feats, scene.points, scene.values,
                               scene.voxel_size)
    torch.testing.assert_close(emb.detach().cpu(), torch.from_numpy(orc), rtol=helpers.RTOL, atol=1e-6)

    g = torch.randn_like(emb)
    emb.backward(g)
    ref.backward(g)
    gv, gx = oracle.trilinear_bwd(vox.cpu().numpy(), xyz.detach().cpu().numpy(), scene.feats, scene.points,
                                  scene.values, scene.voxel_size, g.cpu().numpy())
    helpers.assert_close_scaled(values.grad, gv, what="values.grad vs double-precision oracle")
    helpers.assert_close_scaled(v2.grad, gv, what="torch autograd values.grad vs oracle (sanity)")
    helpers.assert_close_scaled(xyz.grad, gx, what="xyz.grad vs oracle")
    helpers.assert_close_scaled(xyz.grad, x2.grad, what="xyz.grad vs torch autograd")


def test_trilinear_known_answers(cuda):
    """KAT-6: a sample on a corner returns that corner's embedding; the voxel centre returns the mean of 8."""
    scene = synthetic.make_scene("C1")
    pts, feats, values = helpers.scene_tensors(scene, cuda)
    vox = torch.arange(0, scene.n, 7, device=cuda)
    emb_c = ops.trilinear_embed(vox, pts[vox], feats, pts, values, scene.voxel_size)
    torch.testing.assert_close(emb_c, values[feats[vox]].mean(1), rtol=1e-6, atol=1e-7)
    off = torch.tensor([[a, b, c] for a in (-1., 1.) for b in (-1., 1.) for c in (-1., 1.)], device=cuda)
    for j in range(8):
        xyz = pts[vox] + off[j] * (scene.voxel_size / 2)
        emb = ops.trilinear_embed(vox, xyz, feats, pts, values, scene.voxel_size)
        torch.testing.assert_close(emb, values[feats[vox, j]], rtol=1e-5, atol=2e-6)
    assert ops.trilinear_embed(vox[:0], pts[:0], feats, pts, values, scene.voxel_size).shape == (0, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.trilinear_embed(vox.cpu(), pts[vox].cpu(), feats.cpu(), pts.cpu(), values.cpu(), scene.voxel_size)


def _composite_inputs(B, K, seed, cuda):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_valid = torch.randint(0, K + 1, (B,), generator=g)
    mask = torch.arange(K)[None] < n_valid[:, None]
    sigma = torch.randn(B, K, generator=g) * 2
    dists = torch.rand(B, K, generator=g) * 0.05
    fe = (torch.relu(sigma) * dists * 7.0) * mask            # renderer.py:119-121; zero at invalid samples
    tex = torch.rand(B, K, 3, generator=g) * mask[..., None]
    depth = torch.cumsum(torch.rand(B, K, generator=g) * 0.05, 1) + 2.0
    depth = torch.where(mask, depth, torch.full_like(depth, 1e4))   # encoder.py:548
    return fe.to(cuda), tex.to(cuda), depth.to(cuda)


@pytest.mark.parametrize("B,K", [(4096, 127), (1000, 33), (37, 300), (5, 1)])
def test_composite_forward_backward(cuda, B, K):
    fe, tex, depth = _composite_inputs(B, K, 0, cuda)
    fe1, tex1 = fe.clone().requires_grad_(True), tex.clone().requires_grad_(True)
    fe2, tex2 = fe.clone().requires_grad_(True), tex.clone().requires_grad_(True)
    mine = ops.composite(fe1, tex1, depth)
    ref = wrappers.composite_torch(fe2, tex2, depth)
    orc = oracle.composite_fwd(fe.cpu().numpy(), tex.cpu().numpy(), depth.cpu().numpy())
    for a, b, o, nm in zip(mine, ref, orc, ("probs", "depth", "missed", "colors")):
        if nm == "missed":   # missed = 1 - sum(probs) cancels: compare the accumulated opacity that is summed
            a, b, o = 1 - a.detach(), 1 - b.detach(), 1 - torch.from_numpy(o)
        helpers.assert_close_scaled(a, b, what=nm + " vs torch")
        helpers.assert_close_scaled(a, o, what=nm + " vs oracle")
    gs = [torch.randn_like(t) for t in mine]
    torch.autograd.backward(mine, gs)
    torch.autograd.backward(ref, gs)
    g_fe, g_tex = oracle.composite_bwd(fe.cpu().numpy(), tex.cpu().numpy(), depth.cpu().numpy(),
                                       *[g.cpu().numpy() for g in gs])
    helpers.assert_close_scaled(fe1.grad, g_fe, what="d free_energy vs double-precision oracle")
    helpers.assert_close_scaled(tex1.grad, g_tex, what="d texture vs oracle")
    helpers.assert_close_scaled(fe2.grad, g_fe, rtol=1e-4, what="torch autograd d free_energy vs oracle (sanity)")


def test_composite_known_answer(cuda):
    """KAT-4: constant sigma and step: probs_k = (1 - e^{-x}) e^{-x k}, missed = e^{-x K}."""
    B, K, x = 8, 40, 7.0 * 0.8 * 0.03
    fe = torch.full((B, K), x, device=cuda)
    depth = torch.arange(K, device=cuda).float()[None].expand(B, K).contiguous()
    tex = torch.ones(B, K, 3, device=cuda)
    probs, d, missed, colors = ops.composite(fe, tex, depth)
    k = torch.arange(K, device=cuda, dtype=torch.float64)
    want = (1 - np.exp(-x)) * torch.exp(-x * k)
    helpers.assert_close_scaled(probs[0], want, what="KAT-4 probs")
    helpers.assert_close_scaled(1 - missed, torch.full((B,), 1 - float(np.exp(-x * K))), what="KAT-4 1-missed")
    assert float((missed - float(np.exp(-x * K))).abs().max()) < 5e-7     # eps(1.0) level, as in the reference
    helpers.assert_close_scaled(colors, (1 - missed)[:, None].expand(B, 3), what="KAT-4 colors")
    assert ops.composite(fe[:0], tex[:0], depth[:0])[0].shape == (0, K)


def test_fill_in_blend_and_track_voxel_probs(cuda):
    """Per-frame post-processing kernels against the plain-torch statement of the reference code."""
    g = torch.Generator(device="cpu").manual_seed(3)
    N = 10007
    hits = (torch.rand(N, generator=g) < 0.6).to(cuda)
    M = int(hits.sum())
    colors = (torch.rand(M, 3, generator=g) * 2 - 1).to(cuda).requires_grad_(True)
    missed = torch.rand(M, generator=g).to(cuda).requires_grad_(True)
    depths = (torch.rand(M, generator=g) * 4).to(cuda).requires_grad_(True)
    bg = torch.tensor([1.0, 0.5, -1.0], device=cuda)
    mine = ops.fill_in_blend(hits, colors, missed, depths, bg, 5.0)
    c2, m2, d2 = [t.detach().clone().requires_grad_(True) for t in (colors, missed, depths)]
    ref = wrappers.fill_in_blend_torch(hits, c2, m2, d2, bg, 5.0)
    for a, b in zip(mine, ref):
        assert torch.equal(a, b)                       # same mul-then-add per element: bit-identical
    gs = [torch.randn_like(t) for t in mine]
    torch.autograd.backward(mine, gs)
    torch.autograd.backward(ref, gs)
    for a, b in ((colors, c2), (missed, m2), (depths, d2)):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-6, atol=1e-6)
    empty = ops.fill_in_blend(torch.zeros(5, dtype=torch.bool, device=cuda), colors[:0], missed[:0], depths[:0], bg, 5.0)
    assert torch.equal(empty[0], bg.expand(5, 3)) and torch.equal(empty[1], torch.ones(5, device=cuda))

    scene = synthetic.make_scene("C1")
    B, K = 3000, 90
    # a ray meets a (convex) voxel once: its samples in that voxel are consecutive, voxels do not repeat along a ray
    runs = torch.stack([torch.randperm(scene.n, generator=g)[: K // 6] for _ in range(B)]).repeat_interleave(6, dim=1)
    valid = torch.arange(K)[None] < torch.randint(0, K + 1, (B, 1), generator=g)
    idx = torch.where(valid, runs, torch.full_like(runs, -1)).int().to(cuda)
    probs = (torch.rand(B, K, generator=g) * 0.05 * valid).to(cuda)
    start = torch.rand(scene.n, generator=g).to(cuda) * 0.05
    mine_p = ops.track_voxel_probs(start.clone(), idx, probs)
    ref_p = wrappers.track_voxel_probs_torch(start.clone(), idx.long(), probs)
    helpers.assert_close_scaled(mine_p, ref_p, what="max voxel probs")
    assert bool((mine_p >= start).all()) and bool((mine_p > start).any())
