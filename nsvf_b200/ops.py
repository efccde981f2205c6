"""Differentiable fused operators that replace inline torch code of the reference's modules.

  trilinear_embed(sampled_idx, sampled_xyz, feats, centres, values, voxel_size)
        = SparseVoxelEncoder.forward's interpolation, fairnr/modules/encoder.py:582-590
          (trilinear_interp / offset_points, fairnr/data/geometry.py:195-200, 229-238)
  composite(free_energy, texture, sampled_depth)
        = the compositing block of VolumeRenderer.forward_chunk, fairnr/modules/renderer.py:193-218

  linear_layernorm_relu(x, weight, bias, gamma, beta, eps)
        = FCLayer of the field MLP, fairnr/modules/module_utils.py:97-111 (Linear -> LayerNorm -> ReLU): the
          contractions stay on cuBLAS, every other pass of forward and backward is one fused kernel each

All run on the C ABI (include/nsvf_b200.h); CUDA float32 tensors only, no fallback.
"""
import torch
from torch.autograd import Function

from . import _lib, blas

_L = _lib.load()
_p = _lib.ptr


def _need_cuda(**tensors):
    for name, t in tensors.items():
        if t is not None and not t.is_cuda:
            raise RuntimeError("nsvf_b200: %s must be a CUDA tensor (there is no CPU path)" % name)


def as_int32_feats(feats):
    """feats is an int64 buffer in the reference (encoder.py:287); the kernels read int32 keys."""
    return feats if feats.dtype == torch.int32 else feats.to(torch.int32)


class TrilinearEmbed(Function):
    @staticmethod
    def forward(ctx, sampled_idx, sampled_xyz, feats, centres, values, voxel_size):
        _need_cuda(sampled_idx=sampled_idx, sampled_xyz=sampled_xyz, feats=feats, centres=centres, values=values)
        voxel_size = float(voxel_size)
        idx = sampled_idx if (sampled_idx.dtype == torch.int32 and sampled_idx.is_contiguous()) \
            else sampled_idx.to(torch.int32).contiguous()
        xyz = sampled_xyz.detach()
        if xyz.dtype != torch.float32 or not xyz.is_contiguous():
            xyz = xyz.float().contiguous()
        feats32 = feats if (feats.dtype == torch.int32 and feats.is_contiguous()) else as_int32_feats(feats).contiguous()
        centres = centres.detach()
        if centres.dtype != torch.float32 or not centres.is_contiguous():
            centres = centres.float().contiguous()
        vals = values.detach()
        if vals.dtype != torch.float32 or not vals.is_contiguous():
            vals = vals.float().contiguous()
        M, D = idx.numel(), vals.shape[-1]
        out = torch.empty((M, D), dtype=torch.float32, device=vals.device)
        with _lib.device_guard(vals.device):
            _lib.check(_L.nsvf_trilinear_embed_fwd(_lib.current_stream(vals.device), M, D, _p(idx), _p(xyz),
                                                   _p(feats32), _p(centres), _p(vals), voxel_size, _p(out)))
        if ctx.needs_input_grad[4] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(idx, xyz, feats32, centres, vals)
        ctx.voxel_size = voxel_size
        ctx.values_shape = values.shape
        return out if values.dtype == torch.float32 else out.to(values.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        need_values, need_xyz = ctx.needs_input_grad[4], ctx.needs_input_grad[1]
        if not (need_values or need_xyz):
            return None, None, None, None, None, None
        idx, xyz, feats32, centres, vals = ctx.saved_tensors
        M, D = idx.numel(), vals.shape[-1]
        g = grad_out.float().contiguous()
        grad_values = torch.zeros(ctx.values_shape, dtype=torch.float32, device=vals.device)
        grad_xyz = torch.empty((M, 3), dtype=torch.float32, device=vals.device) if need_xyz else None
        with _lib.device_guard(vals.device):
            _lib.check(_L.nsvf_trilinear_embed_bwd(_lib.current_stream(vals.device), M, D, _p(idx), _p(xyz),
                                                   _p(feats32), _p(centres), _p(vals), ctx.voxel_size, _p(g),
                                                   _p(grad_values), _p(grad_xyz)))
        return None, grad_xyz, None, None, (grad_values if need_values else None), None


def trilinear_embed(sampled_idx, sampled_xyz, feats, centres, values, voxel_size):
    """emb[M, D] = sum_j w_j(xyz) * values[feats[idx][j]]  (see include/nsvf_b200.h)."""
    if not torch.is_grad_enabled() or not (values.requires_grad or sampled_xyz.requires_grad):
        # inference: straight to the kernel (an autograd.Function.apply costs ~15 us of host time per call, and the
        # renderer calls this once per window)
        if (sampled_idx.dtype == torch.int32 and sampled_idx.is_contiguous() and sampled_xyz.dtype == torch.float32
                and sampled_xyz.is_contiguous() and feats.dtype == torch.int32 and feats.is_contiguous()
                and centres.dtype == torch.float32 and centres.is_contiguous() and values.dtype == torch.float32
                and values.is_contiguous() and values.is_cuda and sampled_idx.is_cuda):
            M, D = sampled_idx.numel(), values.shape[-1]
            out = torch.empty((M, D), dtype=torch.float32, device=values.device)
            with _lib.device_guard(values.device):
                _lib.check(_L.nsvf_trilinear_embed_fwd(_lib.current_stream(values.device), M, D, _p(sampled_idx),
                                                       _p(sampled_xyz), _p(feats), _p(centres), _p(values),
                                                       float(voxel_size), _p(out)))
            return out
    return TrilinearEmbed.apply(sampled_idx, sampled_xyz, feats, centres, values, voxel_size)


class Composite(Function):
    @staticmethod
    def forward(ctx, free_energy, texture, sampled_depth):
        _need_cuda(free_energy=free_energy, texture=texture, sampled_depth=sampled_depth)
        fe = free_energy.detach().float().contiguous()
        tex = texture.detach().float().contiguous() if texture is not None else None
        dep = sampled_depth.detach().float().contiguous()
        B, K = fe.shape
        dev = fe.device
        probs = torch.empty((B, K), dtype=torch.float32, device=dev)
        depth = torch.empty((B,), dtype=torch.float32, device=dev)
        missed = torch.empty((B,), dtype=torch.float32, device=dev)
        colors = torch.empty((B, 3), dtype=torch.float32, device=dev) if tex is not None else None
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_composite_fwd(_lib.current_stream(dev), B, K, _p(fe), _p(tex), _p(dep), _p(probs),
                                             _p(depth), _p(missed), _p(colors)))
        ctx.save_for_backward(fe, tex, dep)
        ctx.has_tex = tex is not None
        if colors is None:
            colors = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        return probs, depth, missed, colors

    @staticmethod
    def backward(ctx, g_probs, g_depth, g_missed, g_colors):
        fe, tex, dep = ctx.saved_tensors
        B, K = fe.shape
        dev = fe.device
        need_fe, need_tex = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_tex
        if not (need_fe or need_tex):
            return None, None, None

        def c(t):
            return None if t is None else t.float().contiguous()
        g_probs, g_depth, g_missed, g_colors = c(g_probs), c(g_depth), c(g_missed), c(g_colors)
        g_fe = torch.empty((B, K), dtype=torch.float32, device=dev)
        g_tex = torch.empty((B, K, 3), dtype=torch.float32, device=dev) if need_tex else None
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_composite_bwd(_lib.current_stream(dev), B, K, _p(fe), _p(tex), _p(dep), _p(g_probs),
                                             _p(g_depth), _p(g_missed), _p(g_colors if ctx.has_tex else None),
                                             _p(g_fe), _p(g_tex)))
        return (g_fe if need_fe else None), g_tex, None


def composite(free_energy, texture, sampled_depth):
    """(probs[B,K], depth[B], missed[B], colors[B,3]) from free energy, rgb and sample depths."""
    return Composite.apply(free_energy, texture, sampled_depth)


class LinearLayerNormReLU(Function):
    """y = relu(layer_norm(x @ W^T + b) * gamma + beta).  cuBLAS does the three contractions (h, dx, dW); the
    LayerNorm + ReLU forward and {relu mask, d gamma, d beta, dh, d b} of the backward are one kernel each."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, eps):
        _need_cuda(x=x, weight=weight, bias=bias, gamma=gamma, beta=beta)
        if x.dtype != torch.float32 or weight.dtype != torch.float32:
            raise RuntimeError("nsvf_b200: linear_layernorm_relu is float32 only")
        lead = x.shape[:-1]
        x2 = x.detach().reshape(-1, x.shape[-1])
        w, b = weight.detach(), bias.detach()
        g, bt = gamma.detach().contiguous(), beta.detach().contiguous()
        h = torch.addmm(b, x2, w.t())                                  # cuBLAS
        M, N = h.shape
        y = torch.empty_like(h)
        mean = torch.empty(M, dtype=torch.float32, device=h.device)
        rstd = torch.empty(M, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _lib.check(_L.nsvf_ln_relu_fwd(_lib.current_stream(h.device), M, N, _p(h), _p(g), _p(bt), float(eps), _p(y),
                                           _p(mean), _p(rstd)))
        ctx.save_for_backward(x2, w, h, g, bt, mean, rstd)
        ctx.lead = lead
        return y.reshape(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, h, g, bt, mean, rstd = ctx.saved_tensors
        M, N = h.shape
        dev = h.device
        if M == 0:
            z = torch.zeros_like
            return x2.new_zeros(*ctx.lead, x2.shape[-1]), z(w), z(g), z(g), z(g), None
        dy2 = dy.reshape(M, N).float().contiguous()
        LAST_LN_BWD_SHAPE[:] = [M, N]          # read by bench.py: shape of the most recent ln_relu_bwd launch
        dh = torch.empty_like(h)
        sums = torch.empty(3, N, dtype=torch.float32, device=dev)      # d gamma, d beta, d bias
        with torch.cuda.device(dev):
            ws_bytes = _L.nsvf_ln_relu_bwd_workspace_bytes(M, N)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(_L.nsvf_ln_relu_bwd(_lib.current_stream(dev), M, N, _p(h), _p(dy2), _p(g), _p(bt), _p(mean),
                                           _p(rstd), _p(dh), _p(sums[0]), _p(sums[1]), _p(sums[2]), _p(ws), ws_bytes))
        dx = (dh @ w).reshape(*ctx.lead, x2.shape[-1]) if ctx.needs_input_grad[0] else None     # cuBLAS
        dw = _weight_grad(dh, x2) if ctx.needs_input_grad[1] else None                           # cuBLAS
        return dx, dw, sums[2], sums[0], sums[1], None


_DW_SPLIT = 16
LAST_LN_BWD_SHAPE = [0, 0]


def _weight_grad(dh, x2):
    """dW = dh^T @ x, a [N, M] x [M, I] product with M ~ 65536 and a 256 x 416 result.  cuBLAS' BF16x9 path (blas.py)
    has no kernel for that extreme reduction length and falls back to the SIMT SGEMM (0.39 ms); as 16 batched products
    over row blocks plus one sum of the partials it does run on the tensor cores (0.15 ms, and closer to the float64
    result: 4e-7 vs 8e-7 of scale).  The summation order is fixed, so the result is deterministic."""
    M = dh.shape[0]
    if not blas.emulated() or M < 64 * _DW_SPLIT:
        return dh.t() @ x2
    rows = (M // _DW_SPLIT) * _DW_SPLIT
    blk = rows // _DW_SPLIT
    dw = torch.bmm(dh[:rows].reshape(_DW_SPLIT, blk, -1).transpose(1, 2), x2[:rows].reshape(_DW_SPLIT, blk, -1)).sum(0)
    if rows < M:
        dw.addmm_(dh[rows:].t(), x2[rows:])
    return dw


def linear_layernorm_relu(x, weight, bias, gamma, beta, eps=1e-5):
    """FCLayer forward (module_utils.py:97-111) on cuBLAS + the fused LayerNorm/ReLU kernels."""
    return LinearLayerNormReLU.apply(x, weight, bias, gamma, beta, eps)


class PosEnc(Function):
    """NeRFPosEmbLinear(no_linear=True) (module_utils.py:56-87) as one pass; differentiable for the non-angular case."""

    @staticmethod
    def forward(ctx, x, freq, angular, cat_input):
        _need_cuda(x=x, freq=freq)
        if x.dtype != torch.float32:
            raise RuntimeError("nsvf_b200: posenc is float32 only")
        lead, C = x.shape[:-1], x.shape[-1]
        x2 = x.detach().reshape(-1, C).contiguous()
        f = freq.detach().float().contiguous()
        M, L = x2.shape[0], f.numel()
        out = torch.empty(M, C * 2 * L + (C if cat_input else 0), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_L.nsvf_posenc_fwd(_lib.current_stream(x.device), M, C, L, _p(x2), _p(f), int(bool(angular)),
                                          int(bool(cat_input)), _p(out)))
        ctx.save_for_backward(x2, f)
        ctx.meta = (lead, C, bool(angular), bool(cat_input))
        return out.reshape(*lead, out.shape[-1])

    @staticmethod
    def backward(ctx, g):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        x2, f = ctx.saved_tensors
        lead, C, angular, cat_input = ctx.meta
        if angular:
            raise RuntimeError("nsvf_b200: posenc backward is implemented for the non-angular encoding only "
                               "(ray directions carry no gradient)")
        M, L = x2.shape[0], f.numel()
        g2 = g.reshape(M, -1).float().contiguous()
        gx = torch.empty_like(x2)
        with torch.cuda.device(x2.device):
            _lib.check(_L.nsvf_posenc_bwd(_lib.current_stream(x2.device), M, C, L, _p(x2), _p(f), int(cat_input), _p(g2),
                                          _p(gx)))
        return gx.reshape(*lead, C), None, None, None


def posenc(x, freq, angular=False, cat_input=False):
    """[..., C] -> [..., C*2L (+C)]: per channel sin(f_k t) (k < L) then cos(f_k t); t = acos(clamp(x)) if angular."""
    return PosEnc.apply(x, freq, angular, cat_input)


class NarrowLinear(Function):
    """y = x W^T + b for a Linear with 1..4 output features (the sigma / rgb heads): streaming kernels, no GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x=x, weight=weight)
        if x.dtype != torch.float32 or weight.dtype != torch.float32:
            raise RuntimeError("nsvf_b200: narrow_linear is float32 only")
        lead, K = x.shape[:-1], x.shape[-1]
        O = weight.shape[0]
        x2 = x.detach().reshape(-1, K).contiguous()
        w = weight.detach().contiguous()
        b = bias.detach().contiguous() if bias is not None else None
        M = x2.shape[0]
        y = torch.empty(M, O, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_L.nsvf_narrow_linear_fwd(_lib.current_stream(x.device), M, K, O, _p(x2), _p(w), _p(b), _p(y)))
        ctx.save_for_backward(x2, w)
        ctx.meta = (lead, bias is not None)
        return y.reshape(*lead, O)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        lead, has_bias = ctx.meta
        M, K = x2.shape
        O = w.shape[0]
        dev = x2.device
        if M == 0:
            return x2.new_zeros(*lead, K), torch.zeros_like(w), (w.new_zeros(O) if has_bias else None)
        dy2 = dy.reshape(M, O).float().contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(O, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws_bytes = _L.nsvf_narrow_linear_bwd_workspace_bytes(M, K, O)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(_L.nsvf_narrow_linear_bwd(_lib.current_stream(dev), M, K, O, _p(x2), _p(w), _p(dy2), _p(dx), _p(dw),
                                                 _p(db), _p(ws), ws_bytes))
        return (dx.reshape(*lead, K) if dx is not None else None), dw, (db if has_bias else None)


def narrow_linear_supported(in_features, out_features):
    return bool(_L.nsvf_narrow_linear_supported(int(in_features), int(out_features)))


def narrow_linear(x, weight, bias=None):
    return NarrowLinear.apply(x, weight, bias)


class FillInBlend(Function):
    """fill_in x3 + background blend (geometry.py:303-317, nsvf.py:89-104) as one kernel; differentiable w.r.t. the
    compacted colors / missed / depths (the background colour is a constant, `--background-stop-gradient`)."""

    @staticmethod
    def forward(ctx, hits, colors, missed, depths, bg_color, bg_depth):
        _need_cuda(hits=hits, colors=colors, missed=missed, depths=depths)
        dev = colors.device
        N = hits.numel()
        h8 = hits.reshape(-1).to(torch.uint8).contiguous()
        rank = torch.cumsum(h8, 0, dtype=torch.int64)
        bg = torch.as_tensor(bg_color, dtype=torch.float32, device=dev).detach().reshape(3).contiguous()
        out_c = torch.empty((N, 3), dtype=torch.float32, device=dev)
        out_m = torch.empty(N, dtype=torch.float32, device=dev)
        out_d = torch.empty(N, dtype=torch.float32, device=dev)
        c, m, d = colors.detach().float().contiguous(), missed.detach().float().contiguous(), depths.detach().float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_fill_in_blend(_lib.current_stream(dev), N, _p(h8), _p(rank), _p(c), _p(m), _p(d), _p(bg),
                                             float(bg_depth), _p(out_c), _p(out_m), _p(out_d)))
        ctx.save_for_backward(h8, bg)
        ctx.bg_depth = float(bg_depth)
        return out_c, out_m, out_d

    @staticmethod
    def backward(ctx, g_c, g_m, g_d):
        h8, bg = ctx.saved_tensors
        where = h8.bool()
        gc = g_c[where] if g_c is not None else None
        gd = g_d[where] if g_d is not None else None
        gm = g_m[where] if g_m is not None else torch.zeros(int(where.sum()), device=h8.device)
        if gc is not None:
            gm = gm + (gc * bg).sum(-1)
        if gd is not None:
            gm = gm + gd * ctx.bg_depth
        return None, gc, gm, gd, None, None


def fill_in_blend(hits, colors, missed, depths, bg_color, bg_depth):
    """(colors [N,3], missed [N], depths [N]) for all N rays from the results of the hit rays."""
    return FillInBlend.apply(hits, colors, missed, depths, bg_color, bg_depth)


@torch.no_grad()
def track_voxel_probs(max_voxel_probs, voxel_idxs, voxel_probs):
    """In-place max_voxel_probs = max(max_voxel_probs, per-ray sums of voxel_probs per voxel) (encoder.py:594-603)."""
    idx = voxel_idxs.int().contiguous()
    pr = voxel_probs.detach().float().contiguous()
    B, K = idx.shape
    with torch.cuda.device(idx.device):
        _lib.check(_L.nsvf_track_voxel_probs(_lib.current_stream(idx.device), B, K, _p(idx), _p(pr),
                                             max_voxel_probs.numel(), _p(max_voxel_probs)))
    return max_voxel_probs
