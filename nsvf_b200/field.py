"""The radiance-field MLP of the `nsvf_base` architecture (fairnr/models/nsvf.py:168-211 presets;
fairnr/modules/field.py:60-279, implicit.py:41-150, module_utils.py:56-150): 545 297 parameters.

    density  : emb[32] -> posenc(L=6, cat input) = 416 -> 3 x (Linear 256 + LayerNorm + ReLU) = feat[256]
               -> Linear 128 + LayerNorm + ReLU -> Linear 1 = sigma
    texture  : [feat 256, posenc_angular(ray, L=4) = 24] = 280 -> 4 x (Linear 256 + LN + ReLU) -> Linear 3

BASELINE.json's north_star keeps the field's dense contractions on torch/cuBLAS tensor cores (nsvf_b200/blas.py selects
cuBLAS 12.9's fp32-accurate BF16x9 algorithm); every other pass — LayerNorm + ReLU forward/backward with the bias /
gamma / beta reductions, the positional encodings, the 1- and 3-wide output heads — runs in the hand-written kernels of
csrc/field_norm.cu and csrc/field_misc.cu, and GraphedField replays a whole chunk's forward / backward as CUDA graphs.
CUDA only, no fallback: the plain-torch composition of the same network is oracle/field_ref.py (test infrastructure).
"""
import math
import weakref

import torch
import torch.nn as nn

from . import ops


class _PosEnc(nn.Module):
    """NeRFPosEmbLinear(no_linear=True), module_utils.py:56-87, as one hand-written kernel each way
    (csrc/field_misc.cu) instead of outer product + sin + cos + two cats."""

    def __init__(self, in_dim, n_freq, angular, cat_input):
        super().__init__()
        freq = torch.exp(torch.arange(n_freq, dtype=torch.float) * math.log(2.0))
        if not angular:
            freq = freq * math.pi
        self.freq = nn.Parameter(freq, requires_grad=False)
        self.angular, self.cat_input = angular, cat_input
        self.out_dim = in_dim * n_freq * 2 + (in_dim if cat_input else 0)

    def forward(self, x):
        return ops.posenc(x, self.freq, self.angular, self.cat_input)


class _FCLayer(nn.Sequential):
    """FCLayer of the reference (module_utils.py:97-111): Linear -> LayerNorm([o]) -> ReLU.  Parameter names here are the
    flat Sequential's ("0.weight", "0.bias", "1.weight", "1.bias"); the reference nests them under `.net` — reference
    checkpoints load through nsvf_b200.checkpoint, which maps the names both ways.  The contraction runs on cuBLAS, everything else in the hand-written
    kernels of csrc/field_norm.cu: torch's own LayerNorm backward (GammaBetaBackwardCUDAKernel) plus the bias-gradient
    column reductions cost 0.5 ms per call on tall [65536, 256] activations, 40 % of the whole training step.  CUDA
    only, no fallback; the plain-torch composition lives in oracle/field_ref.py (test yardstick, CPU reference arm)."""

    def __init__(self, i, o):
        lin = nn.Linear(i, o)
        nn.init.kaiming_normal_(lin.weight, a=0.0, nonlinearity="relu", mode="fan_in")
        super().__init__(lin, nn.LayerNorm([o]), nn.ReLU())

    def forward(self, x):
        return ops.linear_layernorm_relu(x, self[0].weight, self[0].bias, self[1].weight, self[1].bias, self[1].eps)


class _Head(nn.Linear):
    """Output head (Linear with 1 or 3 output features): streaming kernels of csrc/field_misc.cu — cuBLAS has no
    tensor-core shape for them and spends 130 us on the weight gradient alone."""

    def forward(self, x):
        return ops.narrow_linear(x, self.weight, self.bias)


class RadianceField(nn.Module):
    def __init__(self, embed_dim=32, feat_dim=256, density_dim=128, texture_dim=256, texture_layers=3,
                 feature_layers=1, bg_color=(1.0, 1.0, 1.0), sigma_bias=0.0):
        super().__init__()
        _fc = _FCLayer
        self.emb_enc = _PosEnc(embed_dim, 6, angular=False, cat_input=True)
        self.ray_enc = _PosEnc(3, 4, angular=True, cat_input=False)
        dims = [self.emb_enc.out_dim] + [feat_dim] * (feature_layers + 2)
        self.feature_field = nn.Sequential(*[_fc(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.predictor = nn.Sequential(_fc(feat_dim, density_dim), _Head(density_dim, 1))
        tdims = [feat_dim + self.ray_enc.out_dim] + [texture_dim] * (texture_layers + 1)
        self.renderer = nn.Sequential(*[_fc(a, b) for a, b in zip(tdims[:-1], tdims[1:])], _Head(texture_dim, 3))
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


class _FieldCore(nn.Module):
    """Tensor-in / tensor-out view of a RadianceField (what torch.cuda.make_graphed_callables can capture)."""

    def __init__(self, field):
        super().__init__()
        self.field = field

    def forward(self, emb, ray):
        out = self.field({"emb": emb, "ray": ray})
        return out["sigma"], out["texture"]


class _SlotToken:
    """Lives in the autograd node of one graphed chunk; its finalizer frees the slot if the graph dies without backward."""


class _SlotGuard(torch.autograd.Function):
    """Identity on a replayed chunk's outputs that tells the GraphedField when the chunk's backward has run."""

    @staticmethod
    def forward(ctx, owner_ref, key, token, sigma, texture):
        ctx.owner_ref, ctx.key, ctx.token = owner_ref, key, token
        return sigma.view_as(sigma), texture.view_as(texture)

    @staticmethod
    def backward(ctx, g_sigma, g_texture):
        owner = ctx.owner_ref()
        if owner is not None:
            owner._pending.discard(ctx.key)
        return None, None, None, g_sigma, g_texture


class GraphedField(nn.Module):
    """CUDA-graph replay of the field's forward and backward for the full-size chunks of a training step.

    A training step evaluates the field once per renderer chunk (<= chunk_size = 65536 samples, renderer.py:56): nine
    layers, ~45 kernel launches forward and ~70 backward per chunk, ten chunks per step.  Launched one by one from Python
    that is ~19 ms of host time per step, as long as the kernels themselves take on a B200, so the step is launch-bound.
    Every chunk has the same structure, so it is captured once per chunk slot (torch.cuda.make_graphed_callables: one
    forward and one backward graph per slot, all slots in one memory pool, replayed in capture order fwd 1..k,
    bwd k..1) on inputs padded to `rows`; the outputs are sliced back, so padding rows get zero gradient.
    Chunks that are much smaller than `rows` (the last one of a step), extra chunks beyond `slots`, evaluation, and calls
    that ask for one output only run eagerly on the wrapped field — same kernels, same results."""

    def __init__(self, field, rows=65536, slots=10, min_fill=0.6, eager_first=0):
        super().__init__()
        self.field = field
        self.rows, self.slots, self.min_fill = rows, slots, min_fill
        self.eager_first = eager_first      # bench.py keeps the first chunk of a step eager: its kernels are then
        self._calls = 0                     # launched from the host, where the library's event hook can time them
        self._graphed = None
        self._next = 0
        self._pending = set()               # ids of slot replays whose backward has not run yet
        self._blocked = False
        self._uses = 0
        self.graph_replays = 0
        self.launches_per_replay = 0

    @property
    def bg_color(self):
        return self.field.bg_color

    def begin_step(self):
        """Start of a pipeline forward.  Slots are replayed in capture order from 0 — but only when no slot of an
        earlier forward still waits for its backward: replaying it again would overwrite the static inputs and saved
        activations that backward will read.  Until those backwards have run, chunks evaluate eagerly."""
        self._calls = 0
        self._blocked = len(self._pending) > 0
        if not self._blocked:
            self._next = 0

    def capture(self, dev, embed_dim=32):
        """Capture the chunk graphs.  Must run before any eager forward of the field whose autograd graph is still alive:
        such a graph holds AccumulateGrad nodes bound to the default stream, and the engine's stream hand-over to them
        inside the backward capture would invalidate it (torch warns about exactly this).  bench.py calls it right
        after building the model; otherwise it runs at the first training call."""
        if self._graphed is None:
            self._capture(dev, embed_dim)

    def _capture(self, dev, embed_dim):
        cores = tuple(_FieldCore(self.field) for _ in range(self.slots))
        args = tuple((torch.zeros(self.rows, embed_dim, device=dev, requires_grad=True),
                      torch.nn.functional.normalize(torch.ones(self.rows, 3, device=dev), dim=-1))
                     for _ in range(self.slots))
        from . import _lib
        n0 = _lib.load().nsvf_kernel_launches()
        warmup = 3
        self._graphed = torch.cuda.make_graphed_callables(cores, args, num_warmup_iters=warmup)
        # kernels of this library inside one replayed chunk (forward + backward graph): every slot ran `warmup` eager
        # iterations and one capture, each launching the same kernels
        self.launches_per_replay = (_lib.load().nsvf_kernel_launches() - n0) // (self.slots * (warmup + 1))

    def forward(self, inputs, outputs=("sigma", "texture")):
        emb = inputs.get("emb", None)
        M = 0 if emb is None else emb.shape[0]
        use_graph = (self.training and torch.is_grad_enabled() and emb is not None and emb.is_cuda
                     and inputs.get("feat", None) is None and "sigma" in outputs and "texture" in outputs
                     and emb.requires_grad and emb.dim() == 2
                     and self.rows * self.min_fill <= M <= self.rows and self._next < self.slots
                     and self._calls >= self.eager_first and not self._blocked)
        eager_padded = (self.training and torch.is_grad_enabled() and emb is not None and emb.is_cuda and emb.dim() == 2
                        and inputs.get("feat", None) is None and "sigma" in outputs and "texture" in outputs
                        and self.rows * self.min_fill <= M <= self.rows and self._calls < self.eager_first)
        self._calls += 1
        if self._graphed is None and self.training and torch.is_grad_enabled() and emb is not None and emb.is_cuda:
            self._capture(emb.device, emb.shape[1])       # first training call: before any eager graph exists
        if not use_graph and not eager_padded:
            return self.field(inputs, outputs)
        pad = self.rows - M
        ray = inputs["ray"]
        if pad:
            emb = torch.nn.functional.pad(emb, (0, 0, 0, pad))
            ray = torch.nn.functional.pad(ray, (0, 0, 0, pad))
        if eager_padded:
            # the chunk bench.py keeps eager (so that the library's event hook sees its launches) runs on the same
            # padded shape as the graphed ones: every allocation of the step then has a size the caching allocator has
            # already seen, instead of a new [M, 256] size per step (occasional cudaMalloc + implicit sync)
            out = self.field({"emb": emb, "ray": ray}, outputs)
            inputs["sigma"], inputs["texture"] = out["sigma"][:M], out["texture"][:M]
            return inputs
        sigma, texture = self._graphed[self._next](emb.contiguous(), ray.contiguous())
        self._next += 1
        self.graph_replays += 1
        if sigma.requires_grad:       # this slot must not be replayed again before its backward has consumed it
            self._uses += 1
            key, token = self._uses, _SlotToken()
            self._pending.add(key)
            weakref.finalize(token, self._pending.discard, key)      # graph dropped without a backward
            sigma, texture = _SlotGuard.apply(weakref.ref(self), key, token, sigma, texture)
        inputs["sigma"], inputs["texture"] = sigma[:M], texture[:M]
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
