"""Host-side mirror of the reference's VolumeRenderer (fairnr/modules/renderer.py:51-252) on the sm_100a kernels.

Same interface — forward(input_fn, field_fn, ray_start, ray_dir, samples, encoder_states) -> results dict with
'probs', 'depths', 'max_depths', 'min_depths', 'missed', 'ae', 'colors' — and the SAME chunk schedule
(renderer.py:157-188: field evaluations are issued per group of sample columns whose valid-sample count stays
<= chunk_size, and early termination is evaluated at those chunk boundaries only), so the set of samples that
reach the field is identical to the reference's, including under raymarching_tolerance > 0.

What changes underneath:
  * the schedule is computed from windowed device->host copies of per-column sample counts (48 columns per copy;
    re-fetched after a chunk only when early termination changed who is alive), instead of one `.sum()` host sync
    per sample column (K+1 syncs);
  * boolean-mask compaction of five tensors + masked_scatter back (renderer.py:100,109-114) become the
    compaction kernels (csrc/compact.cu) and one index_put per output;
  * compositing (renderer.py:193-218) is the fused ops.composite kernel with its own backward.
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops

_L = _lib.load()
_p = _lib.ptr


class _ColumnCounts:
    """Lazy per-column valid-sample counts for the chunk scheduler: fetched in windows of `window` columns (one small
    kernel + one D2H copy per window) and invalidated when the early-stop mask changes."""

    def __init__(self, sampled_idx, window=48):
        self.idx, self.window = sampled_idx, window
        self.B, self.K = sampled_idx.shape
        self.es, self.cache, self.lo = None, [], 0

    def set_early_stop(self, early_stop, col0):
        self.es = early_stop.to(torch.uint8).contiguous()
        self.cache, self.lo = [], col0

    def __getitem__(self, k):
        if not (self.lo <= k < self.lo + len(self.cache)):
            lo, hi = k, min(self.K, k + self.window)
            dev = self.idx.device
            counts = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_L.nsvf_masked_col_counts(_lib.current_stream(dev), self.B, self.K, lo, hi, _p(self.idx),
                                                     _p(self.es), _p(counts)))
            self.cache, self.lo = counts.tolist(), lo
        return self.cache[k - self.lo]


def compact_samples(sampled_idx, sampled_depth, sampled_dists, ray_start, ray_dir, col0, col1, early_stop=None,
                    total=None):
    """Row-major compaction of the valid samples of columns [col0, col1).  Returns
    (vox i32 [M], xyz f32 [M,3], dir f32 [M,3], dists f32 [M], flat i64 [M]); `total` (host int) avoids a sync."""
    B, K = sampled_idx.shape
    dev = sampled_idx.device
    # no-ops when forward_chunk already prepared them (it does so once, not once per chunk)
    sampled_idx = sampled_idx.int().contiguous()
    sampled_depth = sampled_depth.float().contiguous()
    sampled_dists = sampled_dists.float().contiguous()
    ray_start, ray_dir = ray_start.float().contiguous(), ray_dir.float().contiguous()
    es = early_stop.to(torch.uint8).contiguous() if early_stop is not None else None
    counts = torch.empty(B, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        _lib.check(_L.nsvf_compact_count(st, B, K, col0, col1, _p(sampled_idx), _p(es), _p(counts)))
        offsets = torch.cumsum(counts, 0)
        M = int(offsets[-1]) if total is None else int(total)
        vox = torch.empty(M, dtype=torch.int32, device=dev)
        xyz = torch.empty((M, 3), dtype=torch.float32, device=dev)
        dirs = torch.empty((M, 3), dtype=torch.float32, device=dev)
        dists = torch.empty(M, dtype=torch.float32, device=dev)
        flat = torch.empty(M, dtype=torch.int64, device=dev)
        if M > 0:
            _lib.check(_L.nsvf_compact_fill(st, B, K, col0, col1, _p(sampled_idx), _p(sampled_depth), _p(sampled_dists),
                                            _p(es), _p(ray_start), _p(ray_dir), _p(offsets), _p(vox), _p(xyz), _p(dirs),
                                            _p(dists), _p(flat)))
    return vox, xyz, dirs, dists, flat


class VolumeRenderer(nn.Module):
    def __init__(self, chunk_size=64, valid_chunk_size=None, discrete_regularization=False,
                 raymarching_tolerance=0.0):
        super().__init__()
        self.chunk_size = 1024 * chunk_size
        self.valid_chunk_size = 1024 * (valid_chunk_size if valid_chunk_size is not None else chunk_size)
        self.discrete_reg = discrete_regularization
        self.raymarching_tolerance = raymarching_tolerance

    # one field evaluation over the valid samples of columns [col0, col1)  (reference forward_once, :77-133)
    def forward_once(self, input_fn, field_fn, ray_start, ray_dir, samples, encoder_states, col0, col1,
                     early_stop=None, total=None, output_types=("sigma", "texture"), noise=None):
        sidx = samples["sampled_point_voxel_idx"]
        vox, xyz, dirs, dists, flat = compact_samples(
            sidx, samples["sampled_point_depth"], samples["sampled_point_distance"], ray_start, ray_dir, col0, col1,
            early_stop, total)
        M = vox.numel()
        if M == 0:
            return None, 0
        field_inputs = input_fn({"sampled_point_voxel_idx": vox, "sampled_point_xyz": xyz,
                                 "sampled_point_ray_direction": dirs, "sampled_point_distance": dists},
                                encoder_states)
        field_outputs = field_fn(field_inputs, outputs=list(output_types))
        out = {"flat": flat}
        if "sigma" in field_outputs:
            sigma = field_outputs["sigma"]
            if noise is None:
                noise = 0 if (not self.discrete_reg and not self.training) else torch.zeros_like(sigma).normal_()
            out["free_energy"] = torch.relu(noise + sigma) * field_inputs["dists"] * 7.0   # renderer.py:117-121
        if "texture" in field_outputs:
            out["texture"] = field_outputs["texture"]
        return out, M

    def forward_chunk(self, input_fn, field_fn, ray_start, ray_dir, samples, encoder_states,
                      output_types=("sigma", "texture"), global_weights=None, noise_fn=None):
        # dense, typed copies ONCE per call (the sampler returns [:, :max_len] views)
        samples = {"sampled_point_depth": samples["sampled_point_depth"].float().contiguous(),
                   "sampled_point_distance": samples["sampled_point_distance"].float().contiguous(),
                   "sampled_point_voxel_idx": samples["sampled_point_voxel_idx"].int().contiguous()}
        ray_start, ray_dir = ray_start.float().contiguous(), ray_dir.float().contiguous()
        sampled_depth = samples["sampled_point_depth"]
        sampled_idx = samples["sampled_point_voxel_idx"]
        B, K = sampled_idx.shape
        dev = sampled_idx.device
        tolerance = self.raymarching_tolerance
        chunk_size = self.chunk_size if self.training else self.valid_chunk_size
        if tolerance > 0:
            tolerance = -math.log(tolerance)
        col_counts = _ColumnCounts(sampled_idx)   # windowed D2H copies (reference: one host sync per column)
        want_tex = "texture" in output_types
        flats, fes, texs = [], [], []           # compacted per-chunk outputs; scattered to [B,K] ONCE at the end
        early_stop, acc_fe, evals = None, None, 0
        size_so_far, start = 0, 0
        for i in range(K + 1):
            if ((i == K) or (size_so_far + col_counts[i] > chunk_size)) and (i > start):
                total = None if early_stop is not None else size_so_far
                out, n = self.forward_once(input_fn, field_fn, ray_start, ray_dir, samples, encoder_states, start, i,
                                           early_stop=early_stop, total=total, output_types=output_types,
                                           noise=None if noise_fn is None else noise_fn(start, i))
                if out is not None:
                    evals += n
                    flats.append(out["flat"])
                    if "free_energy" in out:
                        fes.append(out["free_energy"].float())
                        if tolerance > 0:
                            # per-ray free energy of this chunk (reference: _outputs['free_energy'].sum(1))
                            chunk_fe = torch.zeros(B, dtype=torch.float32, device=dev).index_add_(
                                0, torch.div(out["flat"], K, rounding_mode="floor"), fes[-1].detach())
                            acc_fe = chunk_fe if acc_fe is None else acc_fe + chunk_fe
                            early_stop = acc_fe > tolerance
                            col_counts.set_early_stop(early_stop, i)   # the schedule depends on who stopped
                    if "texture" in out:
                        texs.append(out["texture"].float())
                start, size_so_far = i, 0
            if i < K:
                size_so_far += col_counts[i]

        fe_full = torch.zeros(B * K, dtype=torch.float32, device=dev)
        tex_full = torch.zeros(B * K, 3, dtype=torch.float32, device=dev) if want_tex else None
        if flats:
            flat = torch.cat(flats) if len(flats) > 1 else flats[0]
            if fes:
                fe_full = fe_full.index_put((flat,), torch.cat(fes) if len(fes) > 1 else fes[0])
            if texs:
                tex_full = tex_full.index_put((flat,), torch.cat(texs) if len(texs) > 1 else texs[0])
        fe = fe_full.view(B, K)
        tex = tex_full.view(B, K, 3) if tex_full is not None else None
        hits = sampled_idx.ne(-1)
        if early_stop is not None:
            hits = hits & ~early_stop[:, None]
        probs, depth, missed, colors = ops.composite(fe, tex, sampled_depth)
        if global_weights is not None:   # rarely used; falls back to re-reducing with torch
            probs = probs * global_weights
            depth = (sampled_depth * probs).sum(-1)
            missed = 1 - probs.sum(-1)
            colors = (tex * probs.unsqueeze(-1)).sum(-2) if tex is not None else colors
        results = {
            "probs": probs, "depths": depth,
            "max_depths": sampled_depth.masked_fill(~hits, -1).max(1).values,   # hits after early stop (renderer.py:210)
            "min_depths": sampled_depth.min(1).values,
            "missed": missed, "ae": evals,
        }
        if tex is not None:
            results["colors"] = colors
        return results

    def forward(self, input_fn, field_fn, ray_start, ray_dir, samples, *args, **kwargs):
        chunk_size = self.chunk_size if self.training else self.valid_chunk_size
        if ray_start.size(0) <= chunk_size:
            results = self.forward_chunk(input_fn, field_fn, ray_start, ray_dir, samples, *args, **kwargs)
        else:
            parts = [self.forward_chunk(input_fn, field_fn, ray_start[i: i + chunk_size], ray_dir[i: i + chunk_size],
                                        {name: s[i: i + chunk_size] for name, s in samples.items()}, *args, **kwargs)
                     for i in range(0, ray_start.size(0), chunk_size)]
            results = {name: torch.cat([r[name] for r in parts], 0) if torch.is_tensor(parts[0][name])
                       else sum(r[name] for r in parts) for name in parts[0]}
        if getattr(input_fn, "track_max_probs", False) and (not self.training):
            input_fn.track_voxel_probs(samples["sampled_point_voxel_idx"], results["probs"])
        return results
