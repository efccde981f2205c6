#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref: fairnr's NSVFModel, SparseVoxelEncoder, VolumeRenderer, clib wrappers
and the clib CUDA kernels its own setup.py built for sm_100a) on the GPU of this box, on tensors handed over by bench.py.

BASELINE INFRASTRUCTURE ONLY; runs as a child process of bench.py (the reference flips process-wide torch switches
at import, see ref_loader.py) and never imports the product package.

    python baseline/ref_gpu.py --inputs /tmp/x/inputs.pt --out /tmp/x/ref.json [--legs clib,step,frame]

inputs.pt (written by bench.py): for each config a reference-format state dict (nsvf_b200/checkpoint.py), the rays and
the training targets.  The reference model is built by NSVFModel.build_model from an argparse.Namespace
(nsvf_base defaults, fairnr/models/nsvf.py:168-211) and loaded through its own checkpoint path
(SparseVoxelEncoder.upgrade_state_dict_named + load_state_dict).  Legs:
  clib  : SparseVoxelEncoder.ray_intersect (encoder.py:498-536) -> probs/steps (nsvf.py:65-74) -> ray_sample
          (encoder.py:538-556), i.e. the reference's clib kernels through the reference's own wrappers
  step  : BaseModel._forward (fairnr_model.py:142-186) in train mode with --no-sampling-at-reader, loss, backward,
          Adam — with the reference's RaidanceField and with a contraction-free stand-in field ("hot path only")
  frame : the same _forward in eval mode under no_grad on one 800x800 view, early termination 0.01
All times are CUDA-event times on the current stream after warm-up.  Output tensors of the deterministic frame are
saved next to the JSON so that bench.py can state the parity of the two arms on the very tensors it timed.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.nn as nn

from baseline import ref_loader


class TrivialRefField(nn.Module):
    """The contraction-free stand-in field (same arithmetic as nsvf_b200.field.TrivialField), with the reference
    field's own background module so that NSVFModel.postprocessing works unchanged."""

    def __init__(self, bg_color):
        super().__init__()
        self.bg_color = bg_color

    def forward(self, inputs, outputs=("sigma", "texture")):
        emb = inputs["emb"]
        if "sigma" in outputs:
            inputs["sigma"] = emb[:, 0] * 4 + 1
        if "texture" in outputs:
            inputs["texture"] = torch.tanh(emb[:, 1:4])
        return inputs


def build_model(ns, cfg, dev, train):
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "bbox.txt"), "w") as f:
        f.write(cfg["bbox_line"] + "\n")
    args = argparse.Namespace(
        data=d, initial_boundingbox=os.path.join(d, "bbox.txt"), no_sampling_at_reader=True,
        raymarching_stepsize_ratio=0.125, discrete_regularization=bool(train), min_color=-1,
        transparent_background="1.0,1.0,1.0", background_stop_gradient=True, distributed_rank=0,
        max_hits=int(cfg["max_hits"]), chunk_size=int(cfg["chunk"]), valid_chunk_size=int(cfg["chunk"]),
        raymarching_tolerance=float(cfg["tolerance"]), pixel_per_view=int(cfg.get("pixel_per_view", 2048)),
        use_octree=bool(cfg.get("use_octree", False)))
    ns.nsvf.base_architecture(args)
    model = ns.nsvf.NSVFModel.build_model(args, None)
    sd = {k: v.clone() for k, v in cfg["state"].items()}
    model.encoder.upgrade_state_dict_named(sd, "encoder")       # the reference's own checkpoint resize path
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    model = model.to(dev)
    return model.train(train)


def event_time(fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def leg_clib(ns, cfg, dev, n, warm):
    """intersect (+sort) and sample through the reference encoder; `march` = indices of the rays that are sampled."""
    model = build_model(ns, cfg, dev, train=False)
    enc = model.encoder
    rs, rd = cfg["rs"].to(dev), cfg["rd"].to(dev)
    with torch.no_grad():
        st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=dev))
        t_int, (rs_f, rd_f, inter, hits) = event_time(lambda: enc.ray_intersect(rs, rd, st), n, warm)
        march = cfg.get("march")
        if march is None:
            sel = hits.reshape(-1)
        else:
            sel = torch.zeros_like(hits.reshape(-1))
            sel[march.to(dev)] = True
        sub = {k: v.reshape(-1, v.size(-1))[sel] for k, v in inter.items()}

        def sample():
            o = dict(sub)
            dists = (o["max_depth"] - o["min_depth"]).masked_fill(o["intersected_voxel_idx"].eq(-1), 0)   # nsvf.py:65-74
            o["probs"] = dists / dists.sum(dim=-1, keepdim=True)
            o["steps"] = dists.sum(-1) / enc.step_size
            return enc.ray_sample(o)
        t_smp, samples = event_time(sample, n, warm)
    rays = rd.numel() // 3
    return {"rays_intersected": rays, "rays_sampled": int(sel.sum()), "intersect_ms": round(t_int, 4),
            "sample_ms": round(t_smp, 4), "max_len": int(samples["sampled_point_voxel_idx"].shape[1]),
            "rays_per_s": round(rays / ((t_int + t_smp) / 1e3), 1)}


def leg_step(ns, cfg, dev, n, warm, trivial, anomaly):
    torch.autograd.set_detect_anomaly(anomaly)
    model = build_model(ns, cfg, dev, train=True)
    if trivial:
        model.field = TrivialRefField(model.field.bg_color)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999))
    H, W, V = cfg["H"], cfg["W"], cfg["views"]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    uv = torch.stack([xs.reshape(-1), ys.reshape(-1)], 0)[None, None].expand(1, V, 2, H * W).contiguous().to(dev)
    size = torch.tensor([H, W], dtype=torch.float32, device=dev)[None, None].expand(1, V, 2).contiguous()
    sid = torch.zeros(1, dtype=torch.long, device=dev)
    batches = [tuple(t.to(dev) for t in b) for b in cfg["batches"]]
    state = {"i": 0, "loss": None, "ae": 0}

    def step():
        rs, rd, target = batches[state["i"] % len(batches)]
        state["i"] += 1
        model.set_num_updates(state["i"])
        out = model._forward(rs, rd, uv=uv, size=size, id=sid)
        suv = out["sampled_uv"].reshape(1, V, 2, -1)
        pix = (suv[:, :, 1] * W + suv[:, :, 0]).long() + torch.arange(V, device=dev)[None, :, None] * (H * W)
        sel = target[pix.reshape(-1)]
        rgb = ((out["colors"].reshape(-1, 3) - sel) ** 2).mean() * 128.0
        alpha = (out["missed"].reshape(-1) ** 2).mean()
        loss = rgb + alpha
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        state["loss"], state["ae"] = loss, out["ae"]
        return loss
    ms, loss = event_time(step, n, warm)
    torch.autograd.set_detect_anomaly(False)
    rays = V * int(cfg["pixel_per_view"])
    return {"ms_per_step": round(ms, 3), "rays_per_s": round(rays / (ms / 1e3), 1), "loss": round(float(loss), 5),
            "samples_evaluated": int(state["ae"]), "autograd_anomaly_mode": bool(anomaly)}


def leg_frame(ns, cfg, dev, n, warm, trivial, save=None):
    model = build_model(ns, cfg, dev, train=False)
    if trivial:
        model.field = TrivialRefField(model.field.bg_color)
    rs, rd = cfg["rs"].to(dev), cfg["rd"].to(dev)
    sid = torch.zeros(1, dtype=torch.long, device=dev)

    def frame():
        with torch.no_grad():
            return model._forward(rs, rd, id=sid)
    ms, out = event_time(frame, n, warm)
    if save:
        torch.save({"colors": out["colors"].reshape(-1, 3).cpu(), "depths": out["depths"].reshape(-1).cpu(),
                    "missed": out["missed"].reshape(-1).cpu()}, save)
    return {"ms_per_frame": round(ms, 3), "field_evaluations": int(out["ae"]), "voxels": int(model.encoder.num_voxels)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--legs", default="clib,step,frame")
    a = ap.parse_args()
    t0 = time.time()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    inp = torch.load(a.inputs, weights_only=False)
    ns = ref_loader.load()
    anomaly_default = torch.is_anomaly_enabled()      # reader.py:12 switched it on at import
    torch.autograd.set_detect_anomaly(False)
    res = {"torch": torch.__version__, "ext": os.path.basename(ns.clib._ext.__file__),
           "reference_import_sets_autograd_anomaly_mode": bool(anomaly_default)}
    legs = a.legs.split(",")
    n = int(inp.get("steps", 5))

    def guarded(name, fn):
        try:
            res[name] = fn()
        except Exception as e:      # one failing leg must not lose the others
            res[name] = {"error": repr(e)[:300]}
        print("[ref_gpu] %-28s %s  (%.0f s)" % (name, json.dumps(res[name])[:200], time.time() - t0), file=sys.stderr,
              flush=True)
    if "clib" in legs:
        for c in ("C2", "C3", "C4"):
            if c in inp and "rs" in inp[c]:
                guarded("clib_" + c, lambda c=c: leg_clib(ns, inp[c], dev, 2 if c != "C2" else 3, 1))
    if "step" in legs and "C2" in inp:
        guarded("step_hot_path", lambda: leg_step(ns, inp["C2"], dev, n, 2, trivial=True, anomaly=False))
        guarded("step", lambda: leg_step(ns, inp["C2"], dev, n, 2, trivial=False, anomaly=False))
        guarded("step_stock_anomaly_mode", lambda: leg_step(ns, inp["C2"], dev, max(n // 2, 2), 1, trivial=False,
                                                            anomaly=True))
    if "frame" in legs and "C3" in inp:
        guarded("frame_hot_path", lambda: leg_frame(ns, inp["C3"], dev, 2, 1, trivial=True,
                                                    save=os.path.splitext(a.out)[0] + "_frame.pt"))
        guarded("frame", lambda: leg_frame(ns, inp["C3"], dev, 1, 1, trivial=False))
    res["wall_s"] = round(time.time() - t0, 1)
    with open(a.out, "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
