"""The ray-marching pipeline around the hot path: what BaseModel._forward (fairnr/models/fairnr_model.py:142-186)
with NSVFModel.intersecting / raymarching / postprocessing (fairnr/models/nsvf.py:43-110, nerf.py:49-62) does,
restated compactly on top of the encoder / renderer mirrors.  This is the public call bench.py times:

    pipe = NSVFPipeline(encoder, field, renderer, pixel_per_view=2048)
    out  = pipe(ray_start [S,V,1,3], ray_dir [S,V,P,3])      # S must be 1, like the reference asserts

Training mode with `pixel_per_view > 0` reproduces `--no-sampling-at-reader`: all V*P rays are intersected,
then `pixel_per_view` pixels per view are drawn from the hit mask (Gumbel top-k, reader.py:176-182) and only
those are marched.  Eval mode marches every hit ray (full-frame rendering).
"""
import torch
import torch.nn as nn

TINY = 1e-9


def sampling_without_replacement(logp, k):
    u = torch.rand_like(logp)
    g = -torch.log(-torch.log(u + TINY) + TINY)
    return (logp + g).topk(k, dim=-1)[1]


class NSVFPipeline(nn.Module):
    def __init__(self, encoder, field, renderer, pixel_per_view=0, bg_depth=5.0, hierarchical_sampling=False,
                 fixed_fine_num_samples=0, fine_num_sample_ratio=0.0, field_fine=None):
        super().__init__()
        self.encoder, self.field, self.raymarcher = encoder, field, renderer
        self.field_fine = field_fine
        self.pixel_per_view = pixel_per_view
        self.bg_depth = bg_depth
        self.hierarchical = hierarchical_sampling
        self.fixed_fine_num_samples = fixed_fine_num_samples
        self.fine_num_sample_ratio = fine_num_sample_ratio
        self.padded_samples = False     # True: results["samples"] are the reference's padded [N, max_len] tensors
        # rendering with early termination samples on demand (encoder.ray_sample(lazy=True)): only the per-ray sample
        # counts are computed up front, the renderer then resumes a per-ray serial sampler block by block for the rays
        # that are still alive — work proportional to the evaluated samples (C3 frame: 41 M) instead of the emitted ones
        # (306 M), no row-major sample tensors, no transpose.  Bit-identical results; C3 hot-path frame 13.8 ms against
        # 15.9 ms for the eager sampler + block transpose.  False: the eager sampler.
        self.lazy_sampling = True

    def prepare_hierarchical_sampling(self, inter, samples, results):
        """Bins of the fine pass = the coarse samples (nerf.py:64-79, nsvf.py:83-87)."""
        depth, dists = samples["sampled_point_depth"], samples["sampled_point_distance"]
        out = dict(inter)
        out["min_depth"] = depth - dists * .5
        out["max_depth"] = depth + dists * .5
        out["intersected_voxel_idx"] = samples["sampled_point_voxel_idx"].contiguous()
        safe_probs = results["probs"].detach() + 1e-5
        out["probs"] = safe_probs / safe_probs.sum(-1, keepdim=True)
        steps = safe_probs.new_ones(*safe_probs.size()[:-1])
        if self.fixed_fine_num_samples > 0:
            steps = steps * self.fixed_fine_num_samples
        if self.fine_num_sample_ratio > 0:
            steps = samples["sampled_point_voxel_idx"].ne(-1).sum(-1).float() * self.fine_num_sample_ratio
        out["steps"] = steps
        return out

    def intersecting(self, ray_start, ray_dir, encoder_states):
        S, V, P, _ = ray_dir.size()
        sampled = None
        if self.training and self.pixel_per_view > 0:
            # `--no-sampling-at-reader`: pixels are drawn from the hit mask of ALL rays (nsvf.py:48-60).  The
            # reference materialises [V*P, max_hits] x 3 hit lists for every ray and then gathers the sampled rows;
            # the hit mask alone decides the draw, so we run the any-hit kernel on all rays and the full
            # intersection only on the sampled ones — identical outputs, no [V*P, max_hits] tensors.
            if self.encoder.use_octree:
                ray_start, ray_dir, inter, hits = self.encoder.ray_intersect(ray_start, ray_dir, encoder_states)
            else:
                ray_start, ray_dir, hits = self.encoder.ray_hit_mask(ray_start, ray_dir, encoder_states)
                inter = None
            mask = hits.reshape(S, V, P).float()
            probs = mask / (mask.sum() + 1e-8)
            sampled = sampling_without_replacement(torch.log(probs + TINY), self.pixel_per_view)   # [S,V,k]
            flat = (sampled + torch.arange(V, device=sampled.device)[None, :, None] * P).reshape(S, -1)
            flat, _ = flat.sort(-1)                                     # boolean-mask order of the reference
            take = lambda t: torch.gather(t, 1, flat[..., None].expand(-1, -1, t.size(-1)))
            ray_start, ray_dir = take(ray_start), take(ray_dir)
            if inter is None:
                ray_start, ray_dir, inter, hits = self.encoder.ray_intersect(
                    ray_start.unsqueeze(1), ray_dir.unsqueeze(1), encoder_states)
            else:
                inter = {k: take(v) for k, v in inter.items()}
                hits = torch.gather(hits, 1, flat)
            sampled = flat
        else:
            ray_start, ray_dir, inter, hits = self.encoder.ray_intersect(ray_start, ray_dir, encoder_states)
        min_depth, max_depth, pts_idx = inter["min_depth"], inter["max_depth"], inter["intersected_voxel_idx"]
        # nsvf.py:65-74: dists = (max_depth - min_depth).masked_fill(pts_idx.eq(-1), 0).  Our ray_intersect fills the
        # unused slots with the SAME depth (MAX_DEPTH) in both tensors, so their difference is already +0: the eq() and
        # masked_fill passes over [rays, max_hits] would change nothing
        dists = max_depth - min_depth
        inter["probs"] = dists / dists.sum(dim=-1, keepdim=True)
        inter["steps"] = dists.sum(-1) / self.encoder.step_size
        return ray_start, ray_dir, inter, hits, sampled

    def forward(self, ray_start, ray_dir):
        S, V, P, _ = ray_dir.size()
        assert S == 1, "single object only, like the reference (fairnr_model.py:144)"
        for f in (self.field, self.field_fine):
            if hasattr(f, "begin_step"):      # GraphedField: chunk slots are replayed in capture order within a step
                f.begin_step()
        encoder_states = self.encoder.precompute(id=torch.zeros(1, dtype=torch.long, device=ray_dir.device))
        ray_start, ray_dir, inter, hits, sampled = self.intersecting(ray_start, ray_dir, encoder_states)
        hits_flat = hits.reshape(-1)
        # rays that hit something: ONE nonzero (the only host sync of this stage) + row gathers, where the reference
        # boolean-indexes five tensors (fairnr_model.py:157-159: five nonzero + sync pairs)
        hit_rows = hits_flat.nonzero(as_tuple=True)[0]
        if hit_rows.numel() == hits_flat.numel():       # every ray hit: the gathers would be copies
            inter = {k: v.reshape(-1, *v.shape[2:]) for k, v in inter.items()}
            rs, rd = ray_start.reshape(-1, 3), ray_dir.reshape(-1, 3)
        else:
            inter = {k: v.reshape(-1, *v.shape[2:]).index_select(0, hit_rows) for k, v in inter.items()}
            rs, rd = (ray_start.reshape(-1, 3).index_select(0, hit_rows),
                      ray_dir.reshape(-1, 3).index_select(0, hit_rows))
        encoder_states = {k: v.reshape(-1, v.size(-1)) for k, v in encoder_states.items()}
        dev = ray_dir.device
        results = {"ae": 0}
        bg = self.field.bg_color if hasattr(self.field, "bg_color") else torch.ones(3, device=dev)
        # trimmed sample rows (no padding traffic, no max_len sync) whenever nothing but our renderer reads them
        trimmed = not (self.hierarchical or getattr(self.encoder, "track_max_probs", False) or self.padded_samples)
        if rs.size(0) > 0:
            # with early termination (rendering) the samples are materialised on demand, block by block, for the rays
            # that are still alive; otherwise every sample is needed and the sampler emits them all at once
            lazy = (trimmed and not self.training and getattr(self.raymarcher, "raymarching_tolerance", 0) > 0
                    and self.lazy_sampling)
            samples = self.encoder.ray_sample(inter, trimmed=trimmed, lazy=lazy)
            # 'probs' [B,K] is read by hierarchical sampling and track_voxel_probs only
            need_probs = self.hierarchical or getattr(self.encoder, "track_max_probs", False) or self.padded_samples
            r = self.raymarcher(self.encoder, self.field, rs, rd, samples, encoder_states, return_probs=need_probs)
            if self.hierarchical:      # second, importance-sampled pass (fairnr_model.py:167-173)
                results["coarse"] = {k: r[k] for k in ("colors", "missed", "depths", "probs")}
                inter = self.prepare_hierarchical_sampling(inter, samples, r)
                samples = self.encoder.ray_sample(inter)
                field = self.field_fine if self.field_fine is not None else self.field
                r2 = self.raymarcher(self.encoder, field, rs, rd, samples, encoder_states)
                r2["ae"] = r2["ae"] + r["ae"]
                r = r2
            results["ae"] = r["ae"]
            results["samples"] = samples
            colors, missed, depths = r["colors"], r["missed"], r["depths"]
        else:
            colors, missed, depths = (torch.zeros(0, 3, device=dev), torch.zeros(0, device=dev),
                                      torch.zeros(0, device=dev))
        # fill_in (geometry.py:303-317) + background blend (nsvf.py:89-104), one kernel
        from . import ops
        results["colors"], results["missed"], results["depths"] = ops.fill_in_blend(
            hits_flat, colors, missed, depths, bg, self.bg_depth)
        results["hits"] = hits_flat
        results["sampled"] = sampled
        return results
