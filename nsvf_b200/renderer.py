"""Host-side mirror of the reference's VolumeRenderer (fairnr/modules/renderer.py:51-252) on the sm_100a kernels.

Same interface — forward(input_fn, field_fn, ray_start, ray_dir, samples, encoder_states) -> results dict with
'probs', 'depths', 'max_depths', 'min_depths', 'missed', 'ae', 'colors' — and the SAME chunk schedule
(renderer.py:157-188: field evaluations are issued per group of sample columns whose valid-sample count stays
<= chunk_size, and early termination is evaluated at those chunk boundaries only), so the set of samples that
reach the field is identical to the reference's, including under raymarching_tolerance > 0.

Two implementations of forward_chunk:

  * the ray-marching PLAN (csrc/march.cu; default): the chunk schedule lives on the device.  Samples stay in trimmed
    rows (nothing beyond a ray's last sample is read or written), per window there are three launches of ours
    (single-pass compaction -> [input_fn, field_fn] -> epilogue: free energy, scatter into the rows, early
    termination, next window) and — only with early termination — one 4-int readback; without it the whole window
    list is read once.  Compositing runs on the trimmed rows, and ONE autograd node covers the whole call: its
    backward is the compositing backward followed by one gather kernel per window;
  * the GENERAL path (csrc/compact.cu; inputs whose valid samples are not a prefix of their row, `global_weights`,
    fields without sigma): windowed column counts, count/scan/fill compaction, index_put, dense compositing.
"""
import math
import os

import torch
import torch.nn as nn
from torch.autograd import Function

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


class _MarchRecord:
    """What one forward_chunk of the plan path leaves behind for compositing and its backward."""

    def __init__(self, B, K, ldk, padded_depth_rows, depthT, lens, early_stop, eval_len, feT, texT, lazy=False):
        self.B, self.K, self.ldk, self.lazy = B, K, ldk, lazy
        self.padded_depth_rows, self.depthT = padded_depth_rows, depthT
        self.lens, self.early_stop, self.eval_len, self.feT, self.texT = lens, early_stop, eval_len, feT, texT
        self.windows = []      # (start, end, M, ray_off, sigma, texture, sigma_f32, noise_f32, dists_f32)


class _MarchComposite(Function):
    """Compositing over the slot-major planes a plan run filled (renderer.py:193-218), differentiable w.r.t. every
    window's field outputs: backward = compositing backward + one gather kernel per window (march_epilogue_bwd).
    'probs' is returned as the [B,K] view of its [K,B] plane."""

    @staticmethod
    def forward(ctx, rec, want_probs, *field_outputs):
        B, K = rec.B, rec.K
        dev = rec.feT.device
        ldb = _L.nsvf_march_plane_stride(B)
        probsT = torch.empty((K, ldb), dtype=torch.float32, device=dev) if want_probs else None
        depth = torch.empty(B, dtype=torch.float32, device=dev)
        missed = torch.empty(B, dtype=torch.float32, device=dev)
        colors = torch.empty((B, 3), dtype=torch.float32, device=dev)
        maxd = torch.empty(B, dtype=torch.float32, device=dev)
        mind = torch.empty(B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_march_composite_fwd(
                _lib.current_stream(dev), B, K, _p(rec.eval_len), _p(rec.lens), _p(rec.early_stop), _p(rec.feT),
                _p(rec.texT), _p(rec.depthT), _p(probsT), _p(depth), _p(missed),
                _p(colors if rec.texT is not None else None), _p(maxd), _p(mind), _p(rec.padded_depth_rows), rec.ldk,
                10000.0, int(rec.lazy)))
        if rec.texT is None:
            colors.zero_()
        ctx.rec = rec
        ctx.mark_non_differentiable(maxd, mind)
        if probsT is None:
            probs = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(probs)
        else:
            probs = probsT[:, :B].t()
        return probs, depth, missed, colors, maxd, mind

    @staticmethod
    def backward(ctx, g_probs, g_depth, g_missed, g_colors, _g_maxd, _g_mind):
        rec = ctx.rec
        B, K = rec.B, rec.K
        dev = rec.feT.device

        def c(t):
            return None if t is None else t.float().contiguous()
        ldb = _L.nsvf_march_plane_stride(B)
        g_probsT = None
        if g_probs is not None and g_probs.numel() == B * K and K > 0:
            g_probsT = torch.empty((K, ldb), dtype=torch.float32, device=dev)
            g_probsT[:, :B] = g_probs.t()
        g_depth, g_missed, g_colors = c(g_depth), c(g_missed), c(g_colors)
        has_tex = rec.texT is not None
        g_feT = torch.empty(K * ldb, dtype=torch.float32, device=dev)
        scratch = torch.empty(K * ldb, dtype=torch.float32, device=dev)
        g_texT = torch.empty(K * ldb * 3, dtype=torch.float32, device=dev) if has_tex else None
        grads = []
        with torch.cuda.device(dev):
            st = _lib.current_stream(dev)
            _lib.check(_L.nsvf_march_composite_bwd(
                st, B, K, _p(rec.eval_len), _p(rec.feT), _p(rec.texT), _p(rec.depthT), _p(g_probsT), _p(g_depth),
                _p(g_missed), _p(g_colors if has_tex else None), _p(g_feT), _p(g_texT), _p(scratch)))
            for (start, end, M, ray_off, sigma, texture, sg, nz, dd) in rec.windows:
                gs = torch.empty(M, dtype=torch.float32, device=dev)
                gt = torch.empty((M, 3), dtype=torch.float32, device=dev) if texture is not None else None
                _lib.check(_L.nsvf_march_epilogue_bwd(st, B, K, start, end, _p(ray_off), _p(g_feT), _p(g_texT), _p(sg),
                                                      _p(nz), _p(dd), _p(gs), _p(gt)))
                grads.append(gs.view_as(sigma).to(sigma.dtype))
                if texture is not None:
                    grads.append(gt.view_as(texture).to(texture.dtype))
        return (None, None, *grads)


class VolumeRenderer(nn.Module):
    def __init__(self, chunk_size=64, valid_chunk_size=None, discrete_regularization=False,
                 raymarching_tolerance=0.0):
        super().__init__()
        self.chunk_size = 1024 * chunk_size
        self.valid_chunk_size = 1024 * (valid_chunk_size if valid_chunk_size is not None else chunk_size)
        self.discrete_reg = discrete_regularization
        self.raymarching_tolerance = raymarching_tolerance
        self.return_probs = True      # 'probs' [B,K] of the reference's results; False skips the dense write

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
                      output_types=("sigma", "texture"), global_weights=None, noise_fn=None, return_probs=None):
        """renderer.py:135-232.  `noise_fn(start, end, M)` (tests) supplies the sigma noise of a window instead of
        torch.normal_; `return_probs=False` (extension) leaves results['probs'] empty instead of writing the dense
        [B,K] tensor that only track_voxel_probs and hierarchical sampling read."""
        self._want_probs = self.return_probs if return_probs is None else bool(return_probs)
        lazy = "lazy_pts_idx" in samples
        if lazy or (global_weights is None and "sigma" in output_types
                    and samples["sampled_point_voxel_idx"].dim() == 2 and not os.environ.get("NSVF_RENDER_GENERAL")):
            results = self._forward_chunk_plan(input_fn, field_fn, ray_start, ray_dir, samples, encoder_states,
                                               output_types, noise_fn)
            if results is not None:
                return results
        return self._forward_chunk_general(input_fn, field_fn, ray_start, ray_dir, samples, encoder_states,
                                           output_types, global_weights, noise_fn)

    def _forward_chunk_plan(self, input_fn, field_fn, ray_start, ray_dir, samples, encoder_states, output_types,
                            noise_fn):
        lazy = samples if "lazy_pts_idx" in samples else None
        lens = samples.get("sampled_point_count", None)
        if lazy is not None:
            # on-demand samples: only the per-ray counts exist; blocks of columns are sampled ahead of the window loop
            sidx = depth = dists = None
            B, K, ldk = lens.numel(), int(lazy["lazy_max_steps"]), 0
            dev = lens.device
        else:
            sidx, depth, dists = (samples["sampled_point_voxel_idx"], samples["sampled_point_depth"],
                                  samples["sampled_point_distance"])
            # rows may be strided views (the sampler returns [:, :max_len] slices); anything else is made dense
            ok = (sidx.dtype == torch.int32 and depth.dtype == torch.float32 and dists.dtype == torch.float32
                  and sidx.stride(1) == 1 and depth.stride(1) == 1 and dists.stride(1) == 1
                  and sidx.stride(0) == depth.stride(0) == dists.stride(0))
            if not ok:
                sidx, depth, dists = sidx.int().contiguous(), depth.float().contiguous(), dists.float().contiguous()
            B, K = sidx.shape
            ldk = max(K, sidx.stride(0))
            dev = sidx.device
        ray_start, ray_dir = ray_start.float().contiguous(), ray_dir.float().contiguous()
        tolerance = self.raymarching_tolerance
        chunk_size = self.chunk_size if self.training else self.valid_chunk_size
        tol = -math.log(tolerance) if tolerance > 0 else 0.0
        want_tex = "texture" in output_types
        all_windows = tol <= 0
        rows_padded = lens is None
        grad = torch.is_grad_enabled()
        E = torch.empty
        f32, i32 = torch.float32, torch.int32
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            st = stream.cuda_stream
            plan = torch.zeros(_L.nsvf_march_plan_bytes(B, K) // 4 + 2, dtype=i32, device=dev)
            if lens is None:
                lens = E(B, dtype=i32, device=dev)
                _lib.check(_L.nsvf_march_ray_lengths(st, B, K, ldk, _p(sidx), _p(lens), _p(plan)))
            elif lens.dtype != i32 or not lens.is_contiguous():
                lens = lens.int().contiguous()
            info = self._host_info(16 + 3 * (K + 1))
            info_ptr = info.data_ptr()
            _lib.check(_L.nsvf_march_begin(st, B, K, chunk_size, _p(lens), None, int(all_windows), _p(plan),
                                           info_ptr, info.numel()))
            # slot-major planes of the samples (only the lens[r] valid slots of a row are touched)
            ldb = _L.nsvf_march_plane_stride(B)
            idxT, depthT, distsT = E(K * ldb, dtype=i32, device=dev), E(K * ldb, dtype=f32, device=dev), E(K * ldb, dtype=f32, device=dev)
            early_stop = torch.zeros(B, dtype=torch.uint8, device=dev)
            # without early termination every sample is needed: one transpose.  With it, columns are transposed in
            # blocks just ahead of the window loop, and rays that have stopped are skipped.
            t_block = K if (all_windows or rows_padded) else int(os.environ.get("NSVF_PLANE_BLOCK", 64))
            t_upto = min(K, t_block)
            if lazy is not None:
                lz = (int(lazy["lazy_pts_idx"].size(1)), K, float(lazy["lazy_fixed_step_size"]))
                lz_noise = lazy.get("lazy_noise", None)
                lz_ptrs = (_p(lens), _p(lazy["lazy_quirk"]), _p(lazy["lazy_pts_idx"]), _p(lazy["lazy_min_depth"]),
                           _p(lazy["lazy_max_depth"]), _p(lz_noise), K, 0.5, _p(lazy["lazy_probs"]),
                           _p(lazy["lazy_steps"]), float(lazy["lazy_pad_depth"]))

                # blocks are asked for in increasing order: the resumable serial sampler (one thread per ray, state
                # parked in lz_state between blocks) instead of the any-order block kernel
                lz_state = E(max(_L.nsvf_inverse_cdf_stream_state_bytes(B), 16), dtype=torch.uint8, device=dev)
                p_state = _p(lz_state)

                def fill_planes(k_begin, k_end, stop_ptr):
                    _lib.check(_L.nsvf_inverse_cdf_stream(st, B, lz[0], lz[1], lz[2], k_begin, k_end, stop_ptr, *lz_ptrs,
                                                          p_state, p_idxT, p_depthT, p_distsT))
            else:
                def fill_planes(k_begin, k_end, stop_ptr):
                    _lib.check(_L.nsvf_march_transpose(st, B, K, ldk, k_begin, k_end, stop_ptr, _p(lens), _p(sidx),
                                                       _p(depth), _p(dists), p_idxT, p_depthT, p_distsT))
            p_idxT, p_depthT, p_distsT = _p(idxT), _p(depthT), _p(distsT)
            fill_planes(0, t_upto, None)
            eval_len = torch.zeros(B, dtype=i32, device=dev)
            acc_fe = torch.zeros(B, dtype=f32, device=dev) if tol > 0 else None
            feT = E(K * ldb, dtype=f32, device=dev)
            texT = E(K * ldb * 3, dtype=f32, device=dev) if want_tex else None
            stream.synchronize()
            head = info[:16].tolist()
            if head[4]:
                if lazy is not None:
                    raise RuntimeError("nsvf_b200: lazily sampled rays must have prefix-valid rows")
                return None                      # some row's valid samples are not a prefix: general path
            if all_windows:
                windows = info[16: 16 + 3 * head[5]].view(-1, 3).tolist()
            else:
                windows = [] if head[3] else [head[0:3]]
            record = _MarchRecord(B, K, ldk, depth if rows_padded else None, depthT, lens, early_stop, eval_len, feT,
                                  texT, lazy=t_block < K)
            p_lens, p_es, p_acc, p_ev, p_plan = _p(lens), _p(early_stop), _p(acc_fe), _p(eval_len), _p(plan)
            p_rs, p_rd = _p(ray_start), _p(ray_dir)
            p_feT, p_texT = _p(feT), _p(texT)
            out_types = list(output_types)
            evals, launch_no, w = 0, 0, 0
            ray_off = None
            # With early termination the next window is only known once this window's epilogue has run.  Its compaction
            # is queued right behind that epilogue with the window taken from the plan on the device (start = -1), into
            # buffers of the largest size a window can have, and the host waits on an event recorded BEFORE it: the GPU
            # compacts while the host wakes up, reads the window back and narrows the buffers to its count.
            ahead = (not all_windows) and (not grad) and os.environ.get("NSVF_MARCH_AHEAD", "1") != "0"
            cap_rows = max(int(chunk_size), B)
            spec, held, spare, readback = None, None, [], torch.cuda.Event() if ahead else None
            # our own encoder offers a lean per-window entry for inference (same result as calling it)
            window_fn = getattr(input_fn, "window_fn", None)
            window_fn = window_fn(encoder_states, st) if (window_fn is not None and not grad) else None
            while w < len(windows):
                start, end, M = windows[w]
                w += 1
                if held is not None:
                    spare.append(held)           # the buffers of the window before: free for the one after this
                    held = None
                if spec is not None and spec[4] >= end:
                    vox, xyz, dirs, dists_c = spec[0][:M], spec[1][:M], spec[2][:M], spec[3][:M]
                    held, spec = spec[:4], None
                else:
                    if spec is not None:         # the planes did not reach far enough: compact again, the usual way
                        spare.append(spec[:4])
                        spec = None
                    vox, xyz, dirs = E(M, dtype=i32, device=dev), E((M, 3), dtype=f32, device=dev), E((M, 3), dtype=f32, device=dev)
                    dists_c = E(M, dtype=f32, device=dev)
                    while t_upto < end:
                        t_next = min(K, t_upto + t_block)
                        fill_planes(t_upto, t_next, p_es)
                        t_upto = t_next
                    if grad or ray_off is None:      # the backward needs every window's offsets
                        ray_off = E(B + 1, dtype=i32, device=dev)
                    _lib.check(_L.nsvf_march_compact(st, B, K, start, end, p_lens, p_es, p_idxT, p_depthT, p_distsT, p_rs,
                                                     p_rd, _p(vox), _p(xyz), _p(dirs), _p(dists_c), _p(ray_off), p_plan,
                                                     launch_no))
                    launch_no += 1
                if window_fn is not None:
                    field_inputs = window_fn(vox, xyz, dirs, dists_c)
                else:
                    field_inputs = input_fn({"sampled_point_voxel_idx": vox, "sampled_point_xyz": xyz,
                                             "sampled_point_ray_direction": dirs, "sampled_point_distance": dists_c},
                                            encoder_states)
                field_outputs = field_fn(field_inputs, outputs=out_types)
                sigma = field_outputs["sigma"]
                texture = field_outputs["texture"] if want_tex else None
                if noise_fn is not None:
                    noise = noise_fn(start, end, M)
                elif self.discrete_reg or self.training:
                    noise = torch.zeros_like(sigma).normal_()                       # renderer.py:118
                else:
                    noise = None
                sg = sigma.detach()
                if sg.dtype != f32 or not sg.is_contiguous():
                    sg = sg.float().contiguous()
                tx = None
                if texture is not None:
                    tx = texture.detach()
                    if tx.dtype != f32 or not tx.is_contiguous():
                        tx = tx.float().contiguous()
                nz = noise.float().contiguous() if noise is not None else None
                dd = field_inputs["dists"]
                if dd is not dists_c:
                    dd = dd.detach().float().contiguous()
                _lib.check(_L.nsvf_march_epilogue(st, B, K, start, end, _p(ray_off), p_lens, p_es, p_acc, p_ev,
                                                  _p(sg), _p(nz), _p(dd), _p(tx), tol, p_feT, p_texT, chunk_size,
                                                  int(not all_windows), p_plan, info_ptr))
                evals += M
                if grad:
                    record.windows.append((start, end, M, ray_off, sigma, texture, sg, nz, dd))
                if not all_windows:
                    if ahead:
                        readback.record(stream)
                        want = min(K, end + 2 * (end - start) + 1)       # planes the next window may reach into
                        while t_upto < want:
                            t_next = min(K, t_upto + t_block)
                            fill_planes(t_upto, t_next, p_es)
                            t_upto = t_next
                        bufs = spare.pop() if spare else (
                            E(cap_rows, dtype=i32, device=dev), E((cap_rows, 3), dtype=f32, device=dev),
                            E((cap_rows, 3), dtype=f32, device=dev), E(cap_rows, dtype=f32, device=dev))
                        _lib.check(_L.nsvf_march_compact(st, B, K, -1, -1, p_lens, p_es, p_idxT, p_depthT, p_distsT, p_rs,
                                                         p_rd, _p(bufs[0]), _p(bufs[1]), _p(bufs[2]), _p(bufs[3]),
                                                         _p(ray_off), p_plan, launch_no))
                        launch_no += 1
                        spec = bufs + (t_upto,)
                        readback.synchronize()
                    else:
                        stream.synchronize()
                    head = info[:4].tolist()
                    if not head[3]:
                        windows.append(head[0:3])
            tracked = [t for wnd in record.windows for t in (wnd[4], wnd[5]) if t is not None]
            outs = _MarchComposite.apply(record, self._want_probs, *tracked)
        probs, depth_out, missed, colors, max_depths, min_depths = outs
        results = {"probs": probs, "depths": depth_out, "max_depths": max_depths, "min_depths": min_depths,
                   "missed": missed, "ae": evals}
        if want_tex:
            results["colors"] = colors
        return results

    def _host_info(self, n):
        """Pinned host words the plan kernels publish the schedule to (device writes through the unified address space)."""
        buf = getattr(self, "_info_buf", None)
        if buf is None or buf.numel() < n:
            buf = torch.zeros(max(n, 4096), dtype=torch.int32).pin_memory()
            self._info_buf = buf
        return buf

    def _forward_chunk_general(self, input_fn, field_fn, ray_start, ray_dir, samples, encoder_states,
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
                                           noise=None if noise_fn is None else noise_fn(start, i, None))
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
                                        {name: (s[i: i + chunk_size] if torch.is_tensor(s) and s.dim() > 0 and
                                                s.size(0) == ray_start.size(0) else s) for name, s in samples.items()},
                                        *args, **kwargs)
                     for i in range(0, ray_start.size(0), chunk_size)]
            def merge(name):
                vals = [r[name] for r in parts]
                if not torch.is_tensor(vals[0]):
                    return sum(vals)
                if vals[0].dim() == 2 and vals[0].size(1) > 1 and vals[0].stride(0) == 1:
                    return torch.cat([v.t() for v in vals], 1).t()      # [K,B] planes (probs of the plan path)
                return torch.cat(vals, 0)
            results = {name: merge(name) for name in parts[0]}
        if getattr(input_fn, "track_max_probs", False) and (not self.training):
            input_fn.track_voxel_probs(samples["sampled_point_voxel_idx"], results["probs"])
        return results
