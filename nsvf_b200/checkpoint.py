"""Reference-format checkpoints: state-dict key mapping between NSVFPipeline and the reference's NSVFModel.

The reference nests its field layers as `field.<block>.net.<i>.net.<j>` (FCLayer / ImplicitField / TextureField,
fairnr/modules/module_utils.py:97-111, implicit.py:41-150) and names the density head `predictor.hidden_layer` /
`predictor.output_layer`; our field keeps flat Sequentials.  The encoder buffers (points, keys, feats, num_keys, keep,
voxel_size, step_size, max_hits, values.weight; fairnr/modules/encoder.py:284-311) carry the same names on both sides.
With these two functions a model state moves in either direction, which is also how bench.py hands the SAME voxels
and weights to the unmodified reference (baseline/ref_gpu.py loads them through the reference's own
`upgrade_state_dict_named` + `load_state_dict`, encoder.py:314-346).
"""
import re

import torch

_ENCODER_KEYS = ("points", "keys", "feats", "num_keys", "keep", "voxel_size", "step_size", "max_hits", "values.weight")


def _inner_field(field):
    return getattr(field, "field", field)        # GraphedField wraps the RadianceField


def _field_key_to_reference(key, n_renderer):
    if key == "bg_color":
        return "field.bg_color.bg_color"
    if key == "emb_enc.freq":
        return "field.den_filters.emb.emb"
    if key == "ray_enc.freq":
        return "field.tex_filters.ray.emb"
    m = re.fullmatch(r"feature_field\.(\d+)\.(\d+)\.(\w+)", key)
    if m:
        return "field.feature_field.net.%s.net.%s.%s" % m.groups()
    m = re.fullmatch(r"predictor\.0\.(\d+)\.(\w+)", key)
    if m:
        return "field.predictor.hidden_layer.net.%s.%s" % m.groups()
    m = re.fullmatch(r"predictor\.1\.(\w+)", key)
    if m:
        return "field.predictor.output_layer.%s" % m.group(1)
    m = re.fullmatch(r"renderer\.(\d+)\.(\d+)\.(\w+)", key)
    if m:
        return "field.renderer.net.%s.net.%s.%s" % m.groups()
    m = re.fullmatch(r"renderer\.(\d+)\.(\w+)", key)
    if m and int(m.group(1)) == n_renderer - 1:
        return "field.renderer.net.%s.%s" % m.groups()
    raise KeyError("no reference name for field key %r" % key)


def to_reference_state_dict(pipeline):
    """state_dict of `pipeline` (NSVFPipeline with a RadianceField) under the reference NSVFModel's key names."""
    out = {}
    enc = pipeline.encoder.state_dict()
    for k in _ENCODER_KEYS:
        out["encoder." + k] = enc[k].detach().clone()
    field = _inner_field(pipeline.field)
    if hasattr(field, "renderer"):
        n_renderer = len(field.renderer)
        for k, v in field.state_dict().items():
            out[_field_key_to_reference(k, n_renderer)] = v.detach().clone()
    return out


def load_reference_state_dict(pipeline, state_dict):
    """Load a reference NSVFModel checkpoint (`model` entry of a fairseq checkpoint, or NSVFModel.state_dict()) into
    `pipeline`.  Encoder buffers are resized to the checkpoint's voxel set like the reference does in
    SparseVoxelEncoder.upgrade_state_dict_named (encoder.py:314-346)."""
    enc = pipeline.encoder
    dev = enc.points.device
    for k in _ENCODER_KEYS:
        v = state_dict["encoder." + k].to(dev)
        if k == "values.weight":
            enc.values.weight = torch.nn.Parameter(v.clone().float())
            enc.values.num_embeddings = v.size(0)
        else:
            setattr(enc, k, v.clone())
    enc.clean_runtime_caches()
    field = _inner_field(pipeline.field)
    if hasattr(field, "renderer"):
        n_renderer = len(field.renderer)
        own = field.state_dict()
        mapped = {k: state_dict[_field_key_to_reference(k, n_renderer)] for k in own}
        field.load_state_dict(mapped)
    return pipeline
