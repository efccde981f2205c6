#!/usr/bin/env python
"""Generate golden fixtures from the UNMODIFIED reference Python modules (and the reference build_octree from
oracle/_ref/ref_ext.so), on CPU, in the build container where /root/reference exists.

    python tests/golden/make_cpu_golden.py        # writes tests/golden/cpu_*.npz

The fixtures pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_golden_gpu.py).
Inputs are seeded; every fixture stores its inputs next to the reference outputs, so consumers never need the
reference itself.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_loader  # noqa: E402
import oracle  # noqa: E402
from oracle import wrappers  # noqa: E402
from nsvf_b200 import synthetic  # noqa: E402


def encoder_from_bbox(enc_mod, line, **over):
    d = tempfile.mkdtemp()
    bb = os.path.join(d, "bbox.txt")
    open(bb, "w").write(line + "\n")
    kw = dict(voxel_path=None, initial_boundingbox=bb, voxel_size=None, raymarching_stepsize_ratio=0.125,
              raymarching_stepsize=0.01, max_hits=60, voxel_embed_dim=32, deterministic_step=False, use_octree=False,
              track_max_probs=False)
    kw.update(over)
    return enc_mod.SparseVoxelEncoder(argparse.Namespace(**kw))


def fake_field(inputs, outputs=("sigma", "texture")):
    emb = inputs["emb"]
    if "sigma" in outputs:
        inputs["sigma"] = emb[:, 0] * 6 + emb[:, 5] * 3 + 0.5
    if "texture" in outputs:
        inputs["texture"] = torch.tanh(emb[:, 1:4] * 2)
    return inputs


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    clib, geo, enc_mod, ren_mod, fld = ref_loader.load()
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)

    # ---- KAT-1: encoder buffers from a bbox line -------------------------------------------------------
    e = encoder_from_bbox(enc_mod, "-0.875 -0.875 -0.875 0.875 0.875 0.875 0.25")
    torch.manual_seed(1)
    with torch.no_grad():
        e.values.weight.normal_(0, 32 ** -0.5)
    save("cpu_kat1_encoder.npz", points=e.points.numpy(), feats=e.feats.numpy().astype(np.int32),
         keys=e.keys.numpy().astype(np.int32), step_size=float(e.step_size), max_hits=float(e.max_hits),
         voxel_size=float(e.voxel_size))

    # ---- trilinear interpolation forward / backward (encoder.forward) -----------------------------------
    st = e.precompute(id=None)
    pts, feats, values = st["voxel_center_xyz"], st["voxel_vertex_idx"], st["voxel_vertex_emb"]
    M = 2000
    vox = torch.from_numpy(np.repeat(np.random.randint(0, 512, M // 5), 5)[:M]).long()
    xyz = (pts[vox] + (torch.rand(M, 3) - 0.5) * 0.25).detach().requires_grad_(True)
    out = e.forward({"sampled_point_voxel_idx": vox, "sampled_point_xyz": xyz, "sampled_point_ray_direction": None,
                     "sampled_point_distance": None}, st)
    g = torch.randn(M, 32)
    e.values.weight.grad = None
    out["emb"].backward(g)
    save("cpu_trilinear.npz", points=pts.detach().numpy(), feats=feats.numpy().astype(np.int32),
         values=values.detach().numpy(), voxel_size=float(e.voxel_size), vox=vox.numpy().astype(np.int32),
         xyz=xyz.detach().numpy(), emb=out["emb"].detach().numpy(), grad_out=g.numpy(),
         grad_values=e.values.weight.grad.numpy(), grad_xyz=xyz.grad.numpy())

    # ---- splitting (KAT-2 and a carved set) --------------------------------------------------------------
    for tag, enc in (("full", e), ("carved", None)):
        if enc is None:   # a pruned (carved) set: 7^3 grid with a shell-shaped keep mask
            enc = encoder_from_bbox(enc_mod, "-1.2 -1.2 -1.2 1.2 1.2 1.2 0.4")
            r = enc.points.norm(dim=1)
            enc.keep.copy_(((r > 0.9) & (r < 1.9)).long())
            with torch.no_grad():
                enc.values.weight.normal_(0, 32 ** -0.5)
        st2 = enc.precompute(id=None)
        new_points, new_feats, new_values, new_keys = geo.splitting_points(
            st2["voxel_center_xyz"], st2["voxel_vertex_idx"], st2["voxel_vertex_emb"].detach(), enc.voxel_size / 2.0)
        save("cpu_split_%s.npz" % tag, points=st2["voxel_center_xyz"].numpy(),
             feats=st2["voxel_vertex_idx"].numpy().astype(np.int32), values=st2["voxel_vertex_emb"].detach().numpy(),
             half_voxel=float(enc.voxel_size / 2.0), new_points=new_points.numpy(),
             new_feats=new_feats.numpy().astype(np.int32), new_values=new_values.numpy(),
             new_keys=new_keys.numpy().astype(np.int32))

    # ---- build_octree (reference C++, CPU) on a carved, shuffled lattice -----------------------------------
    from oracle import build_ref
    ref_ext = build_ref.load()
    p0 = synthetic.carve_shell(synthetic.bbox_voxels([-2.4] * 3, [2.4] * 3, 0.4))
    p0 = p0[np.random.RandomState(3).permutation(len(p0))]
    centers, children = geo.build_easy_octree(torch.from_numpy(p0), 0.2)
    coords, residual = geo.discretize_points(torch.from_numpy(p0), 0.2)
    save("cpu_octree.npz", points=p0, half_voxel=0.2, coords=coords.numpy().astype(np.int32),
         centers=centers.numpy(), children=children.numpy().astype(np.int32))

    # ---- renderer: VolumeRenderer.forward_chunk with a fake field, eval, with and without early stop ---------
    scene_pts = st["voxel_center_xyz"].detach().numpy()
    o, d = synthetic.random_rays(700, seed=5)
    idx, dmin, dmax = oracle.aabb_intersect(o[None], d[None], scene_pts, 0.25, 60)
    idx_t, dmin_t, dmax_t, hits = wrappers.sort_hits(*[torch.from_numpy(a[0]) for a in (idx, dmin, dmax)])
    idx_t, dmin_t, dmax_t = idx_t[hits], dmin_t[hits], dmax_t[hits]
    probs, steps = wrappers.probs_and_steps(idx_t, dmin_t, dmax_t, float(e.step_size))
    sidx, sdep, sdist = wrappers.inverse_cdf_sampling(wrappers.NumpyExt(), idx_t, dmin_t, dmax_t, probs, steps, -1, True)
    sidx, sdep, sdist = wrappers.mask_samples(sidx, sdep, sdist)
    rs, rd = torch.from_numpy(o)[hits], torch.from_numpy(d)[hits]
    fix = dict(ray_start=rs.numpy(), ray_dir=rd.numpy(), sampled_idx=sidx.numpy(), sampled_depth=sdep.numpy(),
               sampled_dists=sdist.numpy())
    for tag, tol, chunk in (("plain", 0.0, 64), ("earlystop", 0.05, 4)):
        r = ren_mod.VolumeRenderer(argparse.Namespace(chunk_size=chunk, valid_chunk_size=chunk,
                                                     discrete_regularization=False, raymarching_tolerance=tol,
                                                     trace_normal=False)).eval()
        e.eval()
        samples = {"sampled_point_depth": sdep.clone(), "sampled_point_distance": sdist.clone(),
                   "sampled_point_voxel_idx": sidx.clone()}
        e.values.weight.grad = None
        res = r.forward_chunk(e, fake_field, rs, rd, samples, st)
        loss = (res["colors"] ** 2).sum() + res["missed"].sum() * 0.3 + res["depths"].sum() * 0.1
        loss.backward()
        for k in ("probs", "depths", "missed", "colors", "max_depths", "min_depths"):
            fix["%s_%s" % (tag, k)] = res[k].detach().numpy()
        fix["%s_ae" % tag] = int(res["ae"])
        fix["%s_grad_values" % tag] = e.values.weight.grad.numpy().copy()
    save("cpu_renderer.npz", **fix)

    # ---- pruning with a fake field (keep mask is an integer output) ---------------------------------------
    e2 = encoder_from_bbox(enc_mod, "-0.5 -0.5 -0.5 0.5 0.5 0.5 0.25")   # 5^3 = 125 voxels
    with torch.no_grad():
        e2.values.weight.normal_(0, 32 ** -0.5)
    scores = e2.get_scores(lambda inp, outputs: {"sigma": inp["emb"][:, 0] * 8 - 0.2}, bits=16)
    vals = e2.values.weight.detach().numpy().copy()
    e2.pruning(lambda inp, outputs: {"sigma": inp["emb"][:, 0] * 8 - 0.2}, th=0.5)
    save("cpu_prune.npz", points=e2.points.numpy(), feats=e2.feats.numpy().astype(np.int32), values=vals,
         voxel_size=float(e2.voxel_size), min_score=scores.min(-1)[0].detach().numpy(), keep=e2.keep.numpy().astype(np.int8))
    # ---- the field MLP (RaidanceField, fairnr/modules/field.py:60-279) at the nsvf_base input spec, 128-wide -----
    # pins oracle/field_ref.py (and through it the fused field passes) to the reference's own modules
    make_field_fixture(fld, save)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.startswith("cpu_")))


def field_name_map(ref_key):
    """Reference state_dict key -> key of nsvf_b200.field.RadianceField / oracle.field_ref.ReferenceRadianceField."""
    k = ref_key
    k = k.replace("den_filters.emb.emb", "emb_enc.freq").replace("tex_filters.ray.emb", "ray_enc.freq")
    k = k.replace("bg_color.bg_color", "bg_color")
    k = k.replace("predictor.hidden_layer.net.", "predictor.0.").replace("predictor.output_layer.", "predictor.1.")
    for name in ("feature_field", "renderer"):
        if k.startswith(name + ".net."):
            rest = k[len(name) + 5:]                      # "<i>.net.<j>.weight" or "<i>.weight"
            k = name + "." + rest.replace(".net.", ".")
    return k


def make_field_fixture(field_mod, save, width=128, M=48):
    import types
    args = types.SimpleNamespace(inputs_to_density="emb:6:32", inputs_to_texture="feat:0:%d, ray:4" % width,
                                 feature_embed_dim=width, density_embed_dim=width, texture_embed_dim=width,
                                 feature_layers=1, texture_layers=3, background_stop_gradient=True, min_color=-1,
                                 transparent_background="1.0,1.0,1.0")
    torch.manual_seed(11)
    f = field_mod.RaidanceField(args)
    with torch.no_grad():                                 # non-trivial LayerNorm affine, biases
        for k, p in f.named_parameters():
            if k.endswith("net.1.weight"):
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith("bias") and p.requires_grad:
                p.add_(0.1 * torch.randn_like(p))
    emb = (torch.randn(M, 32) * 0.2).requires_grad_(True)
    ray = torch.nn.functional.normalize(torch.randn(M, 3), dim=-1)
    out = f({"emb": emb, "ray": ray})
    gs, gt = torch.randn(M), torch.randn(M, 3)
    f.zero_grad()
    ((out["sigma"] * gs).sum() + (out["texture"] * gt).sum()).backward()
    fix = {"width": width, "emb": emb.detach().numpy(), "ray": ray.numpy(), "gs": gs.numpy(), "gt": gt.numpy(),
           "sigma": out["sigma"].detach().numpy(), "texture": out["texture"].detach().numpy(),
           "grad_emb": emb.grad.numpy()}
    for k, v in f.state_dict().items():
        fix["w:" + field_name_map(k)] = v.detach().numpy()
    keep = ("feature_field.0.0.weight", "feature_field.0.1.weight", "feature_field.0.1.bias", "feature_field.0.0.bias",
            "predictor.0.0.weight", "predictor.1.weight", "predictor.1.bias", "renderer.0.0.weight", "renderer.4.weight",
            "renderer.4.bias")                                    # a gradient per kind of layer keeps the fixture small
    for k, p in f.named_parameters():
        if p.grad is not None and field_name_map(k) in keep:
            fix["g:" + field_name_map(k)] = p.grad.numpy()
    save("cpu_field.npz", **fix)


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference checkout not found: fixtures can only be regenerated in the build container")
    main()
