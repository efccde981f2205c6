"""GPU parity against golden fixtures produced by the UNMODIFIED reference Python modules
(tests/golden/make_cpu_golden.py): trilinear interpolation fwd/bwd, the VolumeRenderer chunk schedule with and
without early termination, half-voxel splitting (integers bit-exact) and pruning (keep mask bit-exact)."""
import os

import numpy as np
import pytest
import torch

from nsvf_b200 import ops, geometry
from nsvf_b200.encoder import SparseVoxelEncoder
from nsvf_b200.renderer import VolumeRenderer
from tests import helpers

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
load = lambda name: np.load(os.path.join(G, name))
T = lambda a, dev: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_trilinear_matches_reference_encoder_forward(cuda):
    z = load("cpu_trilinear.npz")
    values = T(z["values"], cuda).requires_grad_(True)
    xyz = T(z["xyz"], cuda).requires_grad_(True)
    emb = ops.trilinear_embed(T(z["vox"], cuda), xyz, T(z["feats"], cuda), T(z["points"], cuda), values,
                              float(z["voxel_size"]))
    torch.testing.assert_close(emb.cpu(), torch.from_numpy(z["emb"]), rtol=helpers.RTOL, atol=1e-6)
    emb.backward(T(z["grad_out"], cuda))
    helpers.assert_close_scaled(values.grad, z["grad_values"], what="values.grad vs reference autograd")
    helpers.assert_close_scaled(xyz.grad, z["grad_xyz"], what="xyz.grad vs reference autograd")


def _fake_field(inputs, outputs=("sigma", "texture")):
    emb = inputs["emb"]
    if "sigma" in outputs:
        inputs["sigma"] = emb[:, 0] * 6 + emb[:, 5] * 3 + 0.5
    if "texture" in outputs:
        inputs["texture"] = torch.tanh(emb[:, 1:4] * 2)
    return inputs


@pytest.mark.parametrize("tag,tol,chunk", [("plain", 0.0, 64), ("earlystop", 0.05, 4)])
def test_renderer_matches_reference_forward_chunk(cuda, tag, tol, chunk):
    """Same chunk schedule, same early-termination decisions ('ae' = number of field evaluations is an integer
    output), same composited results and the same gradient into the voxel embeddings."""
    z, k = load("cpu_renderer.npz"), load("cpu_kat1_encoder.npz")
    t = load("cpu_trilinear.npz")
    enc = SparseVoxelEncoder(k["points"], float(k["voxel_size"]), max_hits=60).to(cuda).eval()
    with torch.no_grad():
        enc.values.weight.copy_(T(t["values"], cuda))
    st = enc.precompute(id=None)
    assert torch.equal(st["voxel_center_xyz"].cpu(), torch.from_numpy(t["points"]))
    ren = VolumeRenderer(chunk_size=chunk, valid_chunk_size=chunk, discrete_regularization=False,
                         raymarching_tolerance=tol).eval()
    samples = {"sampled_point_depth": T(z["sampled_depth"], cuda), "sampled_point_distance": T(z["sampled_dists"], cuda),
               "sampled_point_voxel_idx": T(z["sampled_idx"], cuda)}
    res = ren.forward_chunk(enc, _fake_field, T(z["ray_start"], cuda), T(z["ray_dir"], cuda), samples, st)
    assert int(res["ae"]) == int(z[tag + "_ae"]), "different set of field evaluations than the reference"
    for name in ("probs", "depths", "colors", "max_depths", "min_depths"):
        helpers.assert_close_scaled(res[name], z["%s_%s" % (tag, name)], what="%s %s" % (tag, name))
    helpers.assert_close_scaled(1 - res["missed"], 1 - z[tag + "_missed"], what=tag + " 1-missed")
    loss = (res["colors"] ** 2).sum() + res["missed"].sum() * 0.3 + res["depths"].sum() * 0.1
    loss.backward()
    helpers.assert_close_scaled(enc.values.weight.grad, z[tag + "_grad_values"], what=tag + " d loss / d values")


@pytest.mark.parametrize("tag", ["full", "carved"])
def test_splitting_matches_reference(cuda, tag):
    """KAT-2 and a pruned set: new_points and new_feats bit-exact (torch.unique's lexicographic key order),
    Kc' equal, new_values within fp32 tolerance (parent choice is order-undefined in the reference)."""
    z = load("cpu_split_%s.npz" % tag)
    new_points, new_feats, new_values, new_keys = geometry.splitting_points(
        T(z["points"], cuda), T(z["feats"], cuda).long(), T(z["values"], cuda), float(z["half_voxel"]))
    assert torch.equal(new_points.cpu(), torch.from_numpy(z["new_points"]))
    assert new_feats.dtype == torch.int64 and torch.equal(new_feats.cpu().int(), torch.from_numpy(z["new_feats"]))
    assert torch.equal(new_keys.cpu().int(), torch.from_numpy(z["new_keys"]))
    assert new_values.shape == z["new_values"].shape
    torch.testing.assert_close(new_values.cpu(), torch.from_numpy(z["new_values"]), rtol=helpers.RTOL, atol=1e-6)
    if tag == "full":
        assert new_points.shape == (4096, 3) and new_keys.shape == (4913, 3)      # KAT-2: 17^3 keys
    # determinism: two runs give bit-identical values (the reference's scatter_ winner is undefined)
    again = geometry.splitting_points(T(z["points"], cuda), T(z["feats"], cuda).long(), T(z["values"], cuda),
                                      float(z["half_voxel"]))
    assert torch.equal(again[2], new_values) and torch.equal(again[1], new_feats)
    # property: the interpolant is preserved at the old corners (keys at even offsets of the parent)
    assert new_values.shape[0] > z["values"].shape[0]


def test_encoder_splitting_and_pruning_match_reference(cuda):
    z = load("cpu_prune.npz")
    enc = SparseVoxelEncoder(z["points"], float(z["voxel_size"]), max_hits=60).to(cuda)
    assert torch.equal(enc.feats.cpu().int(), torch.from_numpy(z["feats"]))
    with torch.no_grad():
        enc.values.weight.copy_(T(z["values"], cuda))
    field = lambda inp, outputs: {"sigma": inp["emb"][:, 0] * 8 - 0.2}
    scores = enc.get_scores(field, bits=16)
    assert scores.shape == (125, 4096)
    helpers.assert_close_scaled(scores.min(-1)[0], z["min_score"], what="min score per voxel")
    keep, min_score = enc._prune_scores(field, 0.5, bits=16)
    helpers.assert_close_scaled(min_score, z["min_score"], what="fused min score")
    margin = np.abs((1 - z["min_score"]) - 0.5) > 1e-4            # away from the threshold the mask is exact
    assert np.array_equal(keep.cpu().numpy()[margin], z["keep"].astype(bool)[margin]) and margin.mean() > 0.95
    enc.pruning(field, th=0.5)
    assert np.array_equal(enc.keep.cpu().numpy()[margin], z["keep"][margin].astype(np.int64))
    assert 0 < int(enc.keep.sum()) < 125
    n_before = int(enc.keep.sum())
    enc.splitting()                      # splits the kept voxels only; buffers resized like the reference
    assert enc.points.shape[0] == 8 * n_before and enc.feats.shape == (8 * n_before, 8)
    assert int(enc.num_keys) == enc.values.weight.shape[0] and int(enc.keep.sum()) == 8 * n_before
