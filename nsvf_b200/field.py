"""The radiance-field MLP of the `nsvf_base` architecture, on torch / cuBLAS tensor cores.

NOT part of the hand-written hot path: BASELINE.json's north_star keeps the field MLP ("a dense
contraction") on torch/cuBLAS.  It exists so that bench.py can run the named training / rendering step
with random-init weights of the reference architecture (fairnr/models/nsvf.py:168-211 presets;
fairnr/modules/field.py:60-279, implicit.py:41-150, module_utils.py:56-111): 545 297 parameters.

    density  : emb[32] -> posenc(L=6, cat input) = 416 -> 3 x (Linear 256 + LayerNorm + ReLU) = feat[256]
               -> Linear 128 + LayerNorm + ReLU -> Linear 1 = sigma
    texture  : [feat 256, posenc_angular(ray, L=4) = 24] = 280 -> 4 x (Linear 256 + LN + ReLU) -> Linear 3
"""
import math

import torch
import torch.nn as nn


class _PosEnc(nn.Module):
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
        y = y.unsqueeze(-1) * self.freq
        y = torch.cat([torch.sin(y), torch.cos(y)], dim=-1).flatten(-2)
        return torch.cat([y, x], -1) if self.cat_input else y


class _LayerNormReLU(nn.Module):
    """LayerNorm([o]) + ReLU with the same parameters as nn.LayerNorm, written as a non-affine layer_norm followed by
    an explicit scale/shift: torch's fused affine backward (GammaBetaBackwardCUDAKernel) is pathologically slow for
    tall [65536, 256] activations (0.5 ms per call, 30 % of the whole training step); a plain column reduction is not."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = (dim,), eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        return torch.relu(torch.addcmul(self.bias, torch.nn.functional.layer_norm(x, self.dim, None, None, self.eps),
                                        self.weight))


def _fc(i, o):
    lin = nn.Linear(i, o)
    nn.init.kaiming_normal_(lin.weight, a=0.0, nonlinearity="relu", mode="fan_in")
    return nn.Sequential(lin, _LayerNormReLU(o))


class RadianceField(nn.Module):
    def __init__(self, embed_dim=32, feat_dim=256, density_dim=128, texture_dim=256, texture_layers=3,
                 feature_layers=1, bg_color=(1.0, 1.0, 1.0), sigma_bias=0.0):
        super().__init__()
        self.emb_enc = _PosEnc(embed_dim, 6, angular=False, cat_input=True)
        self.ray_enc = _PosEnc(3, 4, angular=True, cat_input=False)
        dims = [self.emb_enc.out_dim] + [feat_dim] * (feature_layers + 2)
        self.feature_field = nn.Sequential(*[_fc(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.predictor = nn.Sequential(_fc(feat_dim, density_dim), nn.Linear(density_dim, 1))
        tdims = [feat_dim + self.ray_enc.out_dim] + [texture_dim] * (texture_layers + 1)
        self.renderer = nn.Sequential(*[_fc(a, b) for a, b in zip(tdims[:-1], tdims[1:])], nn.Linear(texture_dim, 3))
        # transparent_background "1,1,1" with min_color -1 -> b*2-1; background_stop_gradient -> no grad
        self.bg_color = nn.Parameter(torch.tensor([b * 2 - 1 for b in bg_color]), requires_grad=False)
        if sigma_bias:
            with torch.no_grad():
                self.predictor[-1].bias.add_(sigma_bias)

    def forward(self, inputs, outputs=("sigma", "texture")):
        """Same dict-in / dict-out convention as RaidanceField.forward (field.py:217-279)."""
        if inputs.get("feat", None) is None:
            inputs["feat"] = self.feature_field(self.emb_enc(inputs["emb"]))
        if "sigma" in outputs:
            inputs["sigma"] = self.predictor(inputs["feat"]).squeeze(-1)
        if "texture" in outputs:
            inputs["texture"] = self.renderer(torch.cat([inputs["feat"], self.ray_enc(inputs["ray"])], -1))
        return inputs


class TrivialField(nn.Module):
    """Stand-in field without any dense contraction: isolates the hand-written path in benches/tests."""

    def forward(self, inputs, outputs=("sigma", "texture")):
        emb = inputs["emb"]
        if "sigma" in outputs:
            inputs["sigma"] = emb[:, 0] * 4 + 1
        if "texture" in outputs:
            inputs["texture"] = torch.tanh(emb[:, 1:4])
        return inputs
