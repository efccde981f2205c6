"""The nsvf_base field MLP as the reference composes it, in plain PyTorch — TEST INFRASTRUCTURE ONLY.

Restates, module for module, what `fairnr/modules/field.py:60-279` builds for the `nsvf_base` preset
(`fairnr/models/nsvf.py:168-211`) out of `fairnr/modules/module_utils.py`:

  NeRFPosEmbLinear(no_linear=True)   module_utils.py:56-87    -> PosEnc        (outer product, sin, cos, cat, view, cat)
  FCLayer = Linear + LayerNorm + ReLU  module_utils.py:97-111   -> FCLayer
  FCBlock's outermost Linear            module_utils.py:114-150  -> nn.Linear heads

with the same parameter names as nsvf_b200.field.RadianceField, so a state_dict moves between the two.  It is the
yardstick of tests/test_field_gpu.py (the product's fused kernels against this composition) and the MLP of bench.py's CPU
reference arm.  The product (nsvf_b200/) never imports it.
"""
import math

import torch
import torch.nn as nn


class PosEnc(nn.Module):
    def __init__(self, in_dim, n_freq, angular, cat_input):
        super().__init__()
        freq = torch.exp(torch.arange(n_freq, dtype=torch.float) * math.log(2.0))
        if not angular:
            freq = freq * math.pi
        self.freq = nn.Parameter(freq, requires_grad=False)
        self.angular, self.cat_input = angular, cat_input
        self.out_dim = in_dim * n_freq * 2 + (in_dim if cat_input else 0)

    def forward(self, x):
        y = torch.acos(x.clamp(-1 + 1e-6, 1 - 1e-6)) if self.angular else x
        y = y.unsqueeze(-1) * self.freq                    # x.unsqueeze(-1) @ emb.unsqueeze(0): one product per entry
        y = torch.cat([torch.sin(y), torch.cos(y)], dim=-1).flatten(-2)
        return torch.cat([y, x], -1) if self.cat_input else y


class FCLayer(nn.Sequential):
    def __init__(self, i, o):
        lin = nn.Linear(i, o)
        nn.init.kaiming_normal_(lin.weight, a=0.0, nonlinearity="relu", mode="fan_in")
        super().__init__(lin, nn.LayerNorm([o]), nn.ReLU())


class ReferenceRadianceField(nn.Module):
    def __init__(self, embed_dim=32, feat_dim=256, density_dim=128, texture_dim=256, texture_layers=3,
                 feature_layers=1, bg_color=(1.0, 1.0, 1.0), sigma_bias=0.0):
        super().__init__()
        self.emb_enc = PosEnc(embed_dim, 6, angular=False, cat_input=True)
        self.ray_enc = PosEnc(3, 4, angular=True, cat_input=False)
        dims = [self.emb_enc.out_dim] + [feat_dim] * (feature_layers + 2)
        self.feature_field = nn.Sequential(*[FCLayer(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.predictor = nn.Sequential(FCLayer(feat_dim, density_dim), nn.Linear(density_dim, 1))
        tdims = [feat_dim + self.ray_enc.out_dim] + [texture_dim] * (texture_layers + 1)
        self.renderer = nn.Sequential(*[FCLayer(a, b) for a, b in zip(tdims[:-1], tdims[1:])], nn.Linear(texture_dim, 3))
        self.bg_color = nn.Parameter(torch.tensor([b * 2 - 1 for b in bg_color]), requires_grad=False)
        if sigma_bias:
            with torch.no_grad():
                self.predictor[-1].bias.add_(sigma_bias)

    def forward(self, inputs, outputs=("sigma", "texture")):
        if inputs.get("feat", None) is None:
            inputs["feat"] = self.feature_field(self.emb_enc(inputs["emb"]))
        if "sigma" in outputs:
            inputs["sigma"] = self.predictor(inputs["feat"]).squeeze(-1)
        if "texture" in outputs:
            inputs["texture"] = self.renderer(torch.cat([inputs["feat"], self.ray_enc(inputs["ray"])], -1))
        return inputs
