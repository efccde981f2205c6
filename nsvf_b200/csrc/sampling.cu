// Ray sampling for sm_100a: inverse-CDF sampling (the live sampler) and uniform ray sampling.
//
// Replaces fairnr/clib/src/sample_gpu.cu:108-202 (inverse_cdf_sampling_kernel) and :15-106
// (uniform_ray_sampling_kernel) plus the host glue fairnr/clib/src/sample.cpp:23-95.
//
// The per-ray state machine is inherently serial and data dependent, so one lane still owns one ray,
// but (a) the three dense outputs [rays, max_steps] are staged through shared-memory tiles and flushed
// by the whole warp with coalesced stores (the reference writes them with a stride of 4*max_steps
// bytes between lanes), and (b) the kernel itself writes the -1 / 0 padding, which replaces the three
// host-side fill kernels (-ones / zeros) of sample.cpp:40-48, 80-88.
//
// Quirks of the reference that are reproduced on purpose:
//   * `(~done)` is always true, so the trailing loop runs after `done` and emits one more sample whose
//     voxel id is read at pts_idx[H + curr_bin] (curr_bin may equal max_hits: the next ray's slot 0);
//     reads past the end of the tensor are defined here as -1.
//   * the trailing loop's stop test reads pts_idx[curr_bin] of ray 0 of the block row.
//   * a sample that would land at s >= max_steps is dropped (the reference writes out of the row).
#include <cstdlib>

#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

constexpr int kSampWarps = 4;

struct CdfState {
  int curr_bin, s, curr_step, total_steps, phase;  // phase 0: step begin, 1: inside while, 2: trailing, 3: finished
  float curr_min_depth, curr_max_depth, curr_min_cdf, curr_max_cdf, step_size, z_low, curr_cdf;
};

// Emits at most one sample; returns true if one was produced.
__device__ __forceinline__ bool cdf_next(CdfState& st, int max_hits, int max_steps, long long H, int next_idx0,
                                         const int* __restrict__ pts_idx, const int* __restrict__ row0_idx,
                                         const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                                         const float* __restrict__ probs, const float* __restrict__ noise_row,
                                         float noise_const, int& o_idx, float& o_dist, float& o_depth,
                                         int row0_nb = 0 /* leading valid bins of row0 when row0_idx == nullptr */) {
  for (;;) {
    if (st.phase == 0) {
      if (st.curr_step >= st.total_steps) { st.phase = 2; continue; }
      const int ns = st.curr_step < max_steps ? st.curr_step : max_steps - 1;  // reference reads OOB here
      const float nz = noise_row != nullptr ? noise_row[ns] : noise_const;
      st.curr_cdf = __fmul_rn(__fadd_rn((float)st.curr_step, nz), st.step_size);
      st.phase = 1;
    }
    if (st.phase == 1) {
      if (st.curr_cdf > st.curr_max_cdf) {
        o_idx = pts_idx[H + st.curr_bin];
        o_dist = __fsub_rn(st.curr_max_depth, st.z_low);
        o_depth = __fmul_rn(__fadd_rn(st.curr_max_depth, st.z_low), 0.5f);
        st.curr_bin++;
        if (st.curr_bin >= max_hits || pts_idx[H + st.curr_bin] == -1) {
          st.phase = 2;  // done = true; break; if (done) break;
        } else {
          st.curr_min_depth = min_depth[H + st.curr_bin];
          st.curr_max_depth = max_depth[H + st.curr_bin];
          st.curr_min_cdf = st.curr_max_cdf;
          st.curr_max_cdf = __fadd_rn(st.curr_max_cdf, probs[H + st.curr_bin]);
          st.z_low = st.curr_min_depth;
        }
        return true;
      }
      const float u = __fdiv_rn(__fsub_rn(st.curr_cdf, st.curr_min_cdf), __fsub_rn(st.curr_max_cdf, st.curr_min_cdf));
      const float z = __fmaf_rn(u, __fsub_rn(st.curr_max_depth, st.curr_min_depth), st.curr_min_depth);
      o_idx = pts_idx[H + st.curr_bin];
      o_dist = __fsub_rn(z, st.z_low);
      o_depth = __fmul_rn(__fadd_rn(z, st.z_low), 0.5f);
      st.z_low = z;
      st.curr_step++;
      st.phase = 0;
      return true;
    }
    if (st.phase == 2) {
      if (!(st.z_low < st.curr_max_depth)) { st.phase = 3; return false; }
      // reference: pts_idx[H + curr_bin]; curr_bin == max_hits is slot 0 of the next ray in memory
      o_idx = st.curr_bin < max_hits ? pts_idx[H + st.curr_bin] : next_idx0;
      o_dist = __fsub_rn(st.curr_max_depth, st.z_low);
      o_depth = __fmul_rn(__fadd_rn(st.curr_max_depth, st.z_low), 0.5f);
      st.curr_bin++;
      if (st.curr_bin >= max_hits ||
          (row0_idx != nullptr ? row0_idx[st.curr_bin] == -1 : st.curr_bin >= row0_nb)) {
        st.phase = 3;
      } else {
        st.curr_min_depth = min_depth[H + st.curr_bin];
        st.curr_max_depth = max_depth[H + st.curr_bin];
        st.z_low = st.curr_min_depth;
      }
      return true;
    }
    return false;  // phase 3
  }
}

template <int TS>
__global__ void __launch_bounds__(kSampWarps * 32)
inverse_cdf_sampling_kernel(int b, int num_rays, long long valid_rays, int ray_chunk, int max_hits, int max_steps,
                            float fixed_step_size, const int* __restrict__ pts_idx,
                            const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                            const float* __restrict__ noise, float noise_const, const float* __restrict__ probs,
                            const float* __restrict__ steps, int* __restrict__ sampled_idx,
                            float* __restrict__ sampled_depth, float* __restrict__ sampled_dists,
                            int* __restrict__ max_count, int* __restrict__ ray_len, int* __restrict__ holes_flag,
                            float pad_depth, int flags) {
  constexpr int LD = TS + 1;
  __shared__ int t_idx[kSampWarps][32 * LD];
  __shared__ float t_depth[kSampWarps][32 * LD];
  __shared__ float t_dist[kSampWarps][32 * LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_rays = valid_rays;  // rays >= valid_rays are padding aliases of ray 0: never computed
  const long long n_groups = (total_rays + 31) / 32;
  int my_max = 0;

  for (long long g = (long long)blockIdx.x * kSampWarps + warp; g < n_groups; g += (long long)gridDim.x * kSampWarps) {
    const long long ray = g * 32 + lane;
    const bool active = ray < total_rays;
    const long long H = ray * max_hits;
    CdfState st;
    const int* row0_idx = pts_idx;
    const float* noise_row = noise;
    st.phase = 3;
    st.s = 0;
    int next_idx0 = -1, n_valid = 0, lastv = 0, produced = 0;
    if (active) {
      const long long batch = ray / num_rays;
      const int r = (int)(ray - batch * num_rays);
      const int c0 = (r / ray_chunk) * ray_chunk;
      // ray 0 of this block row within its column chunk (reference :194 + clib/__init__.py:259-270)
      const long long row0 = batch * num_rays + c0;
      row0_idx = pts_idx + (row0 < valid_rays ? row0 : 0) * max_hits;
      // the ray that follows this one in the reference's contiguous chunk slice
      long long nxt = -1;
      if (r + 1 < min(c0 + ray_chunk, num_rays)) nxt = ray + 1;
      else if (batch + 1 < b) nxt = (batch + 1) * num_rays + c0;
      if (nxt >= 0) next_idx0 = pts_idx[(nxt < valid_rays ? nxt : 0) * max_hits];
      noise_row = noise != nullptr ? noise + ray * max_steps : nullptr;
      st.curr_bin = 0;
      st.curr_step = 0;
      st.curr_min_depth = min_depth[H];
      st.curr_max_depth = max_depth[H];
      st.curr_min_cdf = 0.0f;
      st.curr_max_cdf = probs[H];
      const float sj = steps[ray];
      st.step_size = __fdiv_rn(1.0f, sj);  // 1.0 / steps[j] in double then rounded: same value
      st.z_low = st.curr_min_depth;
      st.total_steps = min((int)ceilf(sj), max_steps);
      if (fixed_step_size > 0.0f) st.step_size = fixed_step_size;
      st.curr_cdf = 0.0f;
      st.phase = 0;
    }
    const long long out_base = g * 32 * (long long)max_steps;
    const int rows = (int)min((long long)32, total_rays - g * 32);

    for (int t0 = 0; t0 < max_steps; t0 += TS) {
      const int tw = min(TS, max_steps - t0);
      // fill my row of the tile
      int c = 0;
      while (c < tw && st.phase != 3) {
        int oi; float od, oz;
        if (cdf_next(st, max_hits, max_steps, H, next_idx0, pts_idx, row0_idx, min_depth, max_depth, probs,
                     noise_row, noise_const, oi, od, oz)) {
          n_valid += (oi != -1);
          ++produced;
          if (oi != -1) lastv = produced;
          if (flags & 2) {   // ray_sample's post-processing (encoder.py:547-549)
            od = od < 0.f ? 0.f : od;
            if (oi == -1) { od = 0.f; oz = pad_depth; }
          }
          t_idx[warp][lane * LD + c] = oi;
          t_dist[warp][lane * LD + c] = od;
          t_depth[warp][lane * LD + c] = oz;
          ++c;
        }
      }
      for (; c < tw; ++c) {
        t_idx[warp][lane * LD + c] = -1;
        t_dist[warp][lane * LD + c] = 0.0f;
        t_depth[warp][lane * LD + c] = pad_depth;
      }
      __syncwarp();
      // coalesced flush: 32/TS rows per instruction
      constexpr int RPI = 32 / TS;
      const int cc = lane % TS, rsub = lane / TS;
      for (int r = 0; r < rows; r += RPI) {
        const int rr = r + rsub;
        if (rr < rows && cc < tw) {
          const long long o = out_base + (long long)rr * max_steps + t0 + cc;
          sampled_idx[o] = t_idx[warp][rr * LD + cc];
          sampled_depth[o] = t_depth[warp][rr * LD + cc];
          sampled_dists[o] = t_dist[warp][rr * LD + cc];
        }
      }
      __syncwarp();
      // every lane finished: the rest of all rows is padding, written directly
      if (__all_sync(NSVF_FULL_MASK, st.phase == 3)) {
        const int tnext = t0 + TS;
        if (tnext < max_steps && (flags & 1)) {
          const int rem = max_steps - tnext;
          for (int r = 0; r < rows; ++r) {
            const long long o = out_base + (long long)r * max_steps + tnext;
            for (int k = lane; k < rem; k += 32) {
              sampled_idx[o + k] = -1;
              sampled_depth[o + k] = pad_depth;
              sampled_dists[o + k] = 0.0f;
            }
          }
        }
        break;
      }
    }
    my_max = max(my_max, n_valid);
    if (active && ray_len != nullptr) ray_len[ray] = lastv;
    if (active && holes_flag != nullptr && n_valid != lastv) atomicOr(holes_flag, 1);
  }
  if (max_count != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(NSVF_FULL_MASK, my_max, o));
    if (lane == 0 && my_max > 0) atomicMax(max_count, my_max);
  }
}

// ---- warp-per-ray inverse-CDF sampler -------------------------------------------------------------------------
// The reference loop is serial per ray, but its result is a MERGE of two monotone sequences: the step samples
// (cdf_i = (i + noise_i) * step, non-decreasing in i) and the bin boundaries (cumulative probs).  Once the
// cumulative sums exist — computed sequentially, because `cdf > curr_max_cdf` must see the reference's exact
// left-to-right float sums — every step sample is independent:
//   bin b_i   = running max over i' <= i of lower_bound(cum, cdf_i')         (the bin pointer only moves forward)
//   position  = i + b_i                              (b_i bin-end samples were emitted before it)
//   z_low     = z_{i-1} if b_{i-1} == b_i else min_depth[b_i]
// and the bin-end sample of bin j sits at j + #{i : b_i <= j} with z_low = z of the last step sample in bin j (or
// min_depth[j]).  One warp owns one ray: bins are loaded cooperatively (coalesced), 32 steps are resolved per
// iteration, outputs are written nearly contiguously.  The handful of samples of the reference's trailing loop
// (incl. its `(~done)` / next-ray / ray-0 quirks) are replayed serially from the exact state the main loop leaves.
// Rays whose cumulative sums are not monotone / finite (NaN or negative probs) fall back to the serial machine.
__device__ __forceinline__ int warp_incl_max(int x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(NSVF_FULL_MASK, x, o);
    if (lane >= o) x = max(x, y);
  }
  return x;
}
__device__ __forceinline__ int warp_incl_sum(int x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(NSVF_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  return x;
}

constexpr int kCdfWarps = 4;

__global__ void __launch_bounds__(kCdfWarps * 32)
inverse_cdf_warp_kernel(int b, int num_rays, long long valid_rays, int ray_chunk, int P, int max_steps,
                        float fixed_step_size, const int* __restrict__ pts_idx, const float* __restrict__ min_depth,
                        const float* __restrict__ max_depth, const float* __restrict__ noise, float noise_const,
                        const float* __restrict__ probs, const float* __restrict__ steps,
                        int* __restrict__ sampled_idx, float* __restrict__ sampled_depth,
                        float* __restrict__ sampled_dists, int* __restrict__ max_count, int* __restrict__ ray_len,
                        int* __restrict__ holes_flag, float pad_depth, int flags) {
  extern __shared__ __align__(16) float cdf_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = cdf_smem + (size_t)warp * 6 * P;
  int* s_idx = reinterpret_cast<int*>(base);
  float* s_min = base + P;
  float* s_max = base + 2 * P;
  float* s_cum = base + 3 * P;
  int* s_cnt = reinterpret_cast<int*>(base + 4 * P);
  float* s_zlast = base + 5 * P;
  int my_max = 0;

  for (long long ray = (long long)blockIdx.x * kCdfWarps + warp; ray < valid_rays;
       ray += (long long)gridDim.x * kCdfWarps) {
    const long long H = ray * P, K = ray * max_steps;
    const long long batch = ray / num_rays;
    const int r = (int)(ray - batch * num_rays);
    const int c0 = (r / ray_chunk) * ray_chunk;
    const long long row0 = batch * num_rays + c0;
    const int* row0_idx = pts_idx + (row0 < valid_rays ? row0 : 0) * P;
    long long nxt = -1;
    if (r + 1 < min(c0 + ray_chunk, num_rays)) nxt = ray + 1;
    else if (batch + 1 < b) nxt = (batch + 1) * num_rays + c0;
    const int next_idx0 = nxt >= 0 ? pts_idx[(nxt < valid_rays ? nxt : 0) * P] : -1;
    const float* noise_row = noise != nullptr ? noise + ray * max_steps : nullptr;

    // 1) bins -> shared memory (coalesced).  The voxel ids are loaded up to the first -1 (bin 0 is always used): that is
    //    the number of usable bins nb; depths / probabilities are then loaded for bins 0..nb only (a hit list of
    //    max_hits = 135 slots holds ~20 hits: 3 of the 4 row reads shrink to one 128-byte line).  The trailing loop can
    //    look at bins beyond nb when the block row's ray 0 has more hits; it then reads them from global memory.
    int nb = P;
    for (int j0 = 0; j0 < P; j0 += 32) {
      const int j = j0 + lane;
      int v = -1;
      if (j < P) { v = pts_idx[H + j]; s_idx[j] = v; }
      const unsigned m = __ballot_sync(NSVF_FULL_MASK, j >= 1 && j < P && v == -1);
      if (m) {
        nb = j0 + __ffs(m) - 1;
        for (int t = j0 + 32 + lane; t < P; t += 32) s_idx[t] = pts_idx[H + t];   // rare: ids beyond the first -1
        break;
      }
    }
    const int n_ld = min(P, nb + 1);
    bool neg = false;
    for (int j = lane; j < n_ld; j += 32) {
      s_min[j] = min_depth[H + j];
      s_max[j] = max_depth[H + j];
      const float pr = probs[H + j];
      s_cum[j] = pr;
      neg |= (j < nb) && !(pr >= 0.0f);      // negative or NaN probability
      s_cnt[j] = 0;
    }
    const bool nonneg = !__any_sync(NSVF_FULL_MASK, neg);
    __syncwarp();
    // 2) the reference's sequential cumulative sums (left to right, one rounding per add).  With non-negative terms
    //    the sums are non-decreasing (rounding is monotone) and all finite iff the last one is, so the loop carries no
    //    checks: 14 % of this kernel's instructions were the per-element tests of the first version.
    int ok = 1;
    if (lane == 0) {
      float c = s_cum[0];
#pragma unroll 4
      for (int j = 1; j < nb; ++j) {
        c = __fadd_rn(c, s_cum[j]);
        s_cum[j] = c;
      }
      ok = fabsf(c) <= 3.0e38f;
    }
    ok = __shfl_sync(NSVF_FULL_MASK, ok, 0) && nonneg;
    __syncwarp();

    const float sj = steps[ray];
    float step_size = __fdiv_rn(1.0f, sj);
    if (fixed_step_size > 0.0f) step_size = fixed_step_size;
    // steps beyond max_steps could only produce samples at slots >= max_steps, which are dropped anyway
    const int total_steps = min((int)ceilf(sj), max_steps);
    int s_end = 0, n_valid = 0, lastv = 0;
    const bool post = (flags & 2) != 0;   // ray_sample's clamp / masking (encoder.py:547-549) fused in
    auto emit = [&](int pos, int oi, float od, float oz) {
      if (post) {
        od = od < 0.f ? 0.f : od;
        if (oi == -1) { od = 0.f; oz = pad_depth; }
      }
      sampled_idx[K + pos] = oi;
      sampled_dists[K + pos] = od;
      sampled_depth[K + pos] = oz;
      if (oi != -1) { ++n_valid; lastv = max(lastv, pos + 1); }
    };

    if (!ok) {
      // serial fallback (exact reference loop) on lane 0
      if (lane == 0) {
        CdfState st;
        st.curr_bin = 0; st.s = 0; st.curr_step = 0;
        st.curr_min_depth = s_min[0]; st.curr_max_depth = s_max[0];
        st.curr_min_cdf = 0.0f; st.curr_max_cdf = probs[H];
        st.step_size = step_size; st.z_low = st.curr_min_depth; st.total_steps = total_steps;
        st.curr_cdf = 0.0f; st.phase = 0;
        int sidx = 0;
        while (sidx < max_steps && st.phase != 3) {
          int oi; float od, oz;
          if (cdf_next(st, P, max_steps, H, next_idx0, pts_idx, row0_idx, min_depth, max_depth, probs, noise_row,
                       noise_const, oi, od, oz)) {
            emit(sidx, oi, od, oz);
            ++sidx;
          }
        }
        s_end = sidx;
      }
      s_end = __shfl_sync(NSVF_FULL_MASK, s_end, 0);
    } else {
      // 3) step samples, 32 per iteration
      int carry_b = 0, prev_b = -1, n_in = total_steps;
      float prev_z = 0.f;
      bool done = false;
      for (int i0 = 0; i0 < total_steps && !done; i0 += 32) {
        const int i = i0 + lane;
        const bool active = i < total_steps;
        float cdf = 0.f;
        int g = 0;
        if (active) {
          const int ns = i < max_steps ? i : max_steps - 1;
          const float nz = noise_row != nullptr ? noise_row[ns] : noise_const;
          cdf = __fmul_rn(__fadd_rn((float)i, nz), step_size);
        }
        // The bin pointer only moves forward, so every lane's bin lies in [carry_b, w_hi], w_hi = the first bin that is
        // not below the LARGEST cdf of this iteration: one warp max and one ballot over the next 32 bins bound the
        // binary search to the few bins the iteration crosses (~3 probes instead of log2(hits): the full search was
        // 20 % of the kernel's instructions and 27 % of its stall samples).
        int w_hi = nb;
        {
          const float cdf_max = float_from_order_key(__reduce_max_sync(NSVF_FULL_MASK, float_order_key(cdf)));
          const int jb = carry_b + lane;
          const bool stop = jb >= nb || !(cdf_max > s_cum[jb]);
          const unsigned wm = __ballot_sync(NSVF_FULL_MASK, stop);
          if (wm != 0u && cdf_max == cdf_max) w_hi = min(nb, carry_b + __ffs(wm) - 1);
        }
        if (active) {
          int lo = min(carry_b, w_hi), hi = w_hi;   // first j >= carry_b with !(cdf > cum[j]); nb = none
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf > s_cum[mid]) lo = mid + 1; else hi = mid;
          }
          g = lo;
        }
        int bb = max(warp_incl_max(g, lane), carry_b);
        const unsigned dmask = __ballot_sync(NSVF_FULL_MASK, active && bb >= nb);
        int first = 32;
        if (dmask) { first = __ffs(dmask) - 1; n_in = i0 + first; done = true; }
        const bool act = active && lane < first;
        float z = 0.f;
        if (act) {
          const float cmin = bb > 0 ? s_cum[bb - 1] : 0.0f, cmax = s_cum[bb];
          const float u = __fdiv_rn(__fsub_rn(cdf, cmin), __fsub_rn(cmax, cmin));
          z = __fmaf_rn(u, __fsub_rn(s_max[bb], s_min[bb]), s_min[bb]);
        }
        float pz = __shfl_up_sync(NSVF_FULL_MASK, z, 1);
        int pb = __shfl_up_sync(NSVF_FULL_MASK, bb, 1);
        if (lane == 0) { pz = prev_z; pb = prev_b; }
        const int nbn = __shfl_down_sync(NSVF_FULL_MASK, bb, 1);
        const unsigned amask = __ballot_sync(NSVF_FULL_MASK, act);
        if (act) {
          const float zlow = (pb == bb) ? pz : s_min[bb];
          const int pos = i + bb;
          if (pos < max_steps) emit(pos, s_idx[bb], __fsub_rn(z, zlow), __fmul_rn(__fadd_rn(z, zlow), 0.5f));
          atomicAdd(&s_cnt[bb], 1);
          const bool next_act = lane < 31 && ((amask >> (lane + 1)) & 1u);
          if (!next_act || nbn != bb) s_zlast[bb] = z;   // the last step sample of a bin (so far)
        }
        if (amask) {
          const int la = 31 - __clz(amask);
          carry_b = __shfl_sync(NSVF_FULL_MASK, bb, la);
          prev_z = __shfl_sync(NSVF_FULL_MASK, z, la);
          prev_b = carry_b;
        }
        __syncwarp();
      }
      const int b_last = done ? nb : (n_in > 0 ? carry_b : 0);
      // 4) bin-end samples of the bins crossed by the main loop
      int cbase = 0;
      for (int j0 = 0; j0 < b_last; j0 += 32) {
        const int j = j0 + lane;
        const bool a = j < b_last;
        const int cnt = a ? s_cnt[j] : 0;
        const int incl = warp_incl_sum(cnt, lane) + cbase;
        if (a) {
          const int pos = j + incl;
          if (pos < max_steps) {
            const float zlow = cnt > 0 ? s_zlast[j] : s_min[j];
            emit(pos, s_idx[j], __fsub_rn(s_max[j], zlow), __fmul_rn(__fadd_rn(s_max[j], zlow), 0.5f));
          }
        }
        cbase = __shfl_sync(NSVF_FULL_MASK, incl, 31);
      }
      // 5) the reference's trailing loop, replayed from the exact state (uniform scalar code; lane 0 writes)
      int curr_bin, sidx = n_in + b_last;
      float curr_max, zl;
      if (done) {
        curr_bin = nb;
        curr_max = s_max[nb - 1];
        zl = s_cnt[nb - 1] > 0 ? s_zlast[nb - 1] : s_min[nb - 1];
      } else {
        curr_bin = b_last;
        curr_max = s_max[curr_bin];
        zl = n_in > 0 ? prev_z : s_min[0];
      }
      while (zl < curr_max) {
        const int oi = curr_bin < P ? s_idx[curr_bin] : next_idx0;
        if (sidx < max_steps && lane == 0) emit(sidx, oi, __fsub_rn(curr_max, zl), __fmul_rn(__fadd_rn(curr_max, zl), 0.5f));
        ++curr_bin;
        ++sidx;
        if (curr_bin >= P || row0_idx[curr_bin] == -1) break;
        curr_max = curr_bin < n_ld ? s_max[curr_bin] : max_depth[H + curr_bin];
        zl = curr_bin < n_ld ? s_min[curr_bin] : min_depth[H + curr_bin];
      }
      s_end = min(sidx, max_steps);
    }
    // 6) padding (skipped for trimmed rows: consumers then read ray_len samples per row and nothing beyond)
    if (flags & 1) {
      for (int t = s_end + lane; t < max_steps; t += 32) {
        sampled_idx[K + t] = -1;
        sampled_depth[K + t] = pad_depth;
        sampled_dists[K + t] = 0.0f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_valid += __shfl_xor_sync(NSVF_FULL_MASK, n_valid, o);
      lastv = max(lastv, __shfl_xor_sync(NSVF_FULL_MASK, lastv, o));
    }
    my_max = max(my_max, n_valid);
    if (lane == 0) {
      if (ray_len != nullptr) ray_len[ray] = lastv;
      if (holes_flag != nullptr && n_valid != lastv) atomicOr(holes_flag, 1);
    }
    __syncwarp();
  }
  if (max_count != nullptr && lane == 0 && my_max > 0) atomicMax(max_count, my_max);
}

// ---- on-demand inverse-CDF sampling for the ray-marching plan ----------------------------------------------------
// With early termination most samples the reference emits are never evaluated (C3 frame: 305.9 M emitted, 41.3 M
// reach the field).  The sampler is a pure function of (bins, noise): in the merge formulation above, step sample i
// lives at position f(i) = i + b_i with b_i = lower_bound(cum, cdf_i) (cdf_i is non-decreasing in i), and the bin-end
// sample of bin j at j + c_j with c_j = #{i : b_i <= j} = lower_bound over i of (cdf_i > cum[j]).  So
//   * inverse_cdf_plan_kernel computes, per ray and in O(bins + log steps), ONLY the number of samples the reference
//     would emit (ray_len: what the window schedule needs), and
//   * inverse_cdf_block_kernel materialises the samples of a column block [k0, k1) for the rays that are still alive,
//     straight into the slot-major planes of the plan (a CTA takes 32 consecutive rays, stages the block in shared
//     memory and writes 128-byte plane rows) — no row-major sample tensors, no transpose, work ~ evaluated samples.
// Both are bit-identical to the eager kernels on the samples they produce (tests/test_march_gpu.py).  The two
// neighbour-ray quirks of the reference's trailing loop are captured per ray by the plan kernel (next_idx0, and the
// number of valid bins of the block row's ray 0), so the block kernel can work on row slices of the rays.
struct CdfRay {
  int nb, total_steps, n_in, b_last, ok, done;
  float step_size;
};

__device__ __forceinline__ float cdf_at(int i, const float* __restrict__ noise_row, float noise_const, int max_steps,
                                        float step_size) {
  const int ns = i < max_steps ? i : max_steps - 1;
  const float nz = noise_row != nullptr ? noise_row[ns] : noise_const;
  return __fmul_rn(__fadd_rn((float)i, nz), step_size);
}
__device__ __forceinline__ int bin_of(float cdf, const float* __restrict__ s_cum, int nb) {   // first j with !(cdf > cum[j])
  int lo = 0, hi = nb;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf > s_cum[mid]) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ float z_at(float cdf, int bb, const float* __restrict__ s_cum, const float* __restrict__ s_min,
                                      const float* __restrict__ s_max) {
  const float cmin = bb > 0 ? s_cum[bb - 1] : 0.0f, cmax = s_cum[bb];
  const float u = __fdiv_rn(__fsub_rn(cdf, cmin), __fsub_rn(cmax, cmin));
  return __fmaf_rn(u, __fsub_rn(s_max[bb], s_min[bb]), s_min[bb]);
}

// Warp-cooperative: bins of one ray -> shared memory, cumulative sums, step range.  Returns the ray descriptor in all
// lanes.  s_* are the warp's tables of P entries each.
__device__ __forceinline__ CdfRay cdf_ray_setup(long long H, int P, int max_steps, float fixed_step_size, float sj,
                                                const int* __restrict__ pts_idx, const float* __restrict__ min_depth,
                                                const float* __restrict__ max_depth, const float* __restrict__ probs,
                                                const float* __restrict__ noise_row, float noise_const, int* s_idx,
                                                float* s_min, float* s_max, float* s_cum) {
  const int lane = threadIdx.x & 31;
  CdfRay r;
  int nb = P;
  for (int j0 = 0; j0 < P; j0 += 32) {      // valid bins are loaded only up to the first -1 (bin 0 is always used)
    const int j = j0 + lane;
    int v = -1;
    if (j < P) { v = pts_idx[H + j]; s_idx[j] = v; }
    const unsigned m = __ballot_sync(NSVF_FULL_MASK, j >= 1 && j < P && v == -1);
    if (m) { nb = j0 + __ffs(m) - 1; break; }
  }
  bool neg = false;
  for (int j = lane; j < min(P, nb + 1); j += 32) {
    s_min[j] = min_depth[H + j];
    s_max[j] = max_depth[H + j];
    const float pr = probs[H + j];
    s_cum[j] = pr;
    neg |= (j < nb) && !(pr >= 0.0f);
  }
  for (int j = nb + 1 + lane; j < P; j += 32) s_idx[j] = -1;     // the trailing loop may look at bins beyond nb
  const bool nonneg = !__any_sync(NSVF_FULL_MASK, neg);
  __syncwarp();
  int ok = 1;
  if (lane == 0) {   // non-negative terms: non-decreasing sums, all finite iff the last one is (see the eager kernel)
    float c = s_cum[0];
#pragma unroll 4
    for (int j = 1; j < nb; ++j) {
      c = __fadd_rn(c, s_cum[j]);
      s_cum[j] = c;
    }
    ok = fabsf(c) <= 3.0e38f;
  }
  r.ok = __shfl_sync(NSVF_FULL_MASK, ok, 0) && nonneg;
  __syncwarp();
  r.nb = nb;
  r.step_size = fixed_step_size > 0.0f ? fixed_step_size : __fdiv_rn(1.0f, sj);
  r.total_steps = min((int)ceilf(sj), max_steps);
  // the main loop stops at the first step whose bin is >= nb (cdf beyond the last cumulative sum)
  int n_in = r.total_steps, done = 0, b_last = 0;
  if (r.ok && lane == 0) {
    if (r.total_steps > 0 && cdf_at(r.total_steps - 1, noise_row, noise_const, max_steps, r.step_size) > s_cum[nb - 1]) {
      int lo = 0, hi = r.total_steps - 1;          // first i with cdf_i > cum[nb-1]
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf_at(mid, noise_row, noise_const, max_steps, r.step_size) > s_cum[nb - 1]) hi = mid; else lo = mid + 1;
      }
      n_in = lo;
      done = 1;
    }
    b_last = done ? nb : (n_in > 0 ? bin_of(cdf_at(n_in - 1, noise_row, noise_const, max_steps, r.step_size), s_cum, nb) : 0);
  }
  r.n_in = __shfl_sync(NSVF_FULL_MASK, n_in, 0);
  r.done = __shfl_sync(NSVF_FULL_MASK, done, 0);
  r.b_last = __shfl_sync(NSVF_FULL_MASK, b_last, 0);
  return r;
}

// The reference's trailing loop (sample_gpu.cu:187-200) replayed from the state the main loop leaves; calls
// emit(position, idx, dist, depth) for every sample it produces (lane-uniform code).  row0_nb = number of leading
// valid bins of ray 0 of the block row (its stop test reads THAT ray's idx), next_idx0 = slot 0 of the next ray.
template <typename Emit>
__device__ __forceinline__ void cdf_trailing(const CdfRay& r, int P, const float* __restrict__ noise_row,
                                             float noise_const, int max_steps, int row0_nb, int next_idx0,
                                             const int* s_idx, const float* s_min, const float* s_max,
                                             const float* s_cum, const float* __restrict__ g_min,
                                             const float* __restrict__ g_max, Emit emit) {
  // bins beyond nb are not staged (cdf_ray_setup loads nb + 1 of them): the loop reads them from global memory
  auto bmin = [&](int j) { return j <= r.nb ? s_min[j] : g_min[j]; };
  auto bmax = [&](int j) { return j <= r.nb ? s_max[j] : g_max[j]; };
  int curr_bin, sidx = r.n_in + r.b_last;
  float curr_max, zl;
  // z of the last step sample, and whether it fell into the bin the trailing loop starts from
  float z_prev = 0.f;
  int b_prev = -1;
  if (r.n_in > 0) {
    const float c = cdf_at(r.n_in - 1, noise_row, noise_const, max_steps, r.step_size);
    b_prev = bin_of(c, s_cum, r.nb);
    z_prev = z_at(c, b_prev, s_cum, s_min, s_max);
  }
  if (r.done) {
    curr_bin = r.nb;
    curr_max = s_max[r.nb - 1];
    zl = b_prev == r.nb - 1 ? z_prev : s_min[r.nb - 1];
  } else {
    curr_bin = r.b_last;
    curr_max = bmax(curr_bin);
    zl = r.n_in > 0 ? z_prev : s_min[0];
  }
  while (zl < curr_max) {
    const int oi = curr_bin < P ? s_idx[curr_bin] : next_idx0;
    emit(sidx, oi, __fsub_rn(curr_max, zl), __fmul_rn(__fadd_rn(curr_max, zl), 0.5f));
    ++curr_bin;
    ++sidx;
    if (curr_bin >= P || curr_bin >= row0_nb) break;
    curr_max = bmax(curr_bin);
    zl = bmin(curr_bin);
  }
}

// ---- resumable serial sampler: the samples of positions [k0, k1), one thread per ray, straight into the planes ------
// The block kernel below can produce any block in any order, at the price of rebuilding a ray's bin tables and locating
// the block inside the merge for every block.  The ray-marching loop asks for blocks in increasing order, and then the
// reference's own serial machine (cdf_next: one sample per call, no searches, no tables) is the cheapest producer there
// is: its 11-word state is parked in global memory between blocks, lanes are consecutive rays, and since every call
// emits exactly one sample all 32 lanes write the SAME plane row — 128 contiguous bytes per store, no staging tile, no
// transpose.  Work is proportional to the samples actually requested (C3 frame: 41 M evaluated of 306 M).
struct __align__(16) CdfParked {
  int curr_bin, curr_step, total_steps, phase;
  float curr_min_depth, curr_max_depth, curr_min_cdf, curr_max_cdf;
  float step_size, z_low, curr_cdf, pad;
};
static_assert(sizeof(CdfParked) == 48, "CdfParked is 48 bytes");
constexpr int kStreamThreads = 128;

__global__ void __launch_bounds__(kStreamThreads)
inverse_cdf_stream_kernel(long long B, long long ldb, int P, int max_steps, float fixed_step_size, int k0, int k1,
                          const unsigned char* __restrict__ early_stop, const int* __restrict__ ray_len,
                          const int2* __restrict__ quirk, const int* __restrict__ pts_idx,
                          const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                          const float* __restrict__ noise, long long noise_stride, float noise_const,
                          const float* __restrict__ probs, const float* __restrict__ steps, float pad_depth,
                          CdfParked* __restrict__ parked, int* __restrict__ idxT, float* __restrict__ depthT,
                          float* __restrict__ distsT) {
  const long long ray = (long long)blockIdx.x * kStreamThreads + threadIdx.x;
  if (ray >= B) return;
  if (early_stop != nullptr && early_stop[ray] != 0) return;
  const int hi = min(ray_len[ray], k1);
  if (hi <= k0) return;
  const long long H = ray * P;
  const int2 qk = quirk[ray];
  const int row0_nb = qk.y >= 0 ? qk.y : -1 - qk.y;
  const float* noise_row = noise != nullptr ? noise + ray * noise_stride : nullptr;
  CdfState st;
  st.s = 0;
  if (k0 == 0) {
    st.curr_bin = 0;
    st.curr_step = 0;
    st.curr_min_depth = min_depth[H];
    st.curr_max_depth = max_depth[H];
    st.curr_min_cdf = 0.0f;
    st.curr_max_cdf = probs[H];
    const float sj = steps[ray];
    st.step_size = fixed_step_size > 0.0f ? fixed_step_size : __fdiv_rn(1.0f, sj);
    st.z_low = st.curr_min_depth;
    st.total_steps = min((int)ceilf(sj), max_steps);
    st.curr_cdf = 0.0f;
    st.phase = 0;
  } else {
    const int4 a = reinterpret_cast<const int4*>(parked + ray)[0];
    const float4 b = reinterpret_cast<const float4*>(parked + ray)[1], c = reinterpret_cast<const float4*>(parked + ray)[2];
    st.curr_bin = a.x; st.curr_step = a.y; st.total_steps = a.z; st.phase = a.w;
    st.curr_min_depth = b.x; st.curr_max_depth = b.y; st.curr_min_cdf = b.z; st.curr_max_cdf = b.w;
    st.step_size = c.x; st.z_low = c.y; st.curr_cdf = c.z;
  }
  int pos = k0;
  while (pos < hi && st.phase != 3) {
    int oi;
    float od, oz;
    if (cdf_next(st, P, max_steps, H, qk.x, pts_idx, nullptr, min_depth, max_depth, probs, noise_row, noise_const, oi, od,
                 oz, row0_nb)) {
      od = od < 0.f ? 0.f : od;                     // ray_sample's clamp / masking (encoder.py:547-549)
      if (oi == -1) { od = 0.f; oz = pad_depth; }
      const long long o = (long long)pos * ldb + ray;
      idxT[o] = oi;
      depthT[o] = oz;
      distsT[o] = od;
      ++pos;
    }
  }
  reinterpret_cast<int4*>(parked + ray)[0] = make_int4(st.curr_bin, st.curr_step, st.total_steps, st.phase);
  reinterpret_cast<float4*>(parked + ray)[1] = make_float4(st.curr_min_depth, st.curr_max_depth, st.curr_min_cdf, st.curr_max_cdf);
  reinterpret_cast<float4*>(parked + ray)[2] = make_float4(st.step_size, st.z_low, st.curr_cdf, 0.f);
}

constexpr int kLazyWarps = 8;       // block kernel: 32 rays per CTA, 4 per warp
constexpr int kLazyBlock = 64;      // positions per block (multiple of 32)

// One THREAD per ray.  Only the number of samples is wanted, and that needs three things from the bins: the total
// probability (is there a step beyond it?), the bin of the last in-range step with the depth it produces, and the
// few bins the trailing loop walks — two sequential passes over the ray's probabilities in registers, no tables.  The
// first generation spent one warp (and 4 P floats of shared memory) per ray on the full set-up of the eager kernel:
// 1.18 ms for the 640 k rays of the C3 frame, latency-bound on lane 0's dependent adds.
constexpr int kPlanThreads = 128;

__global__ void __launch_bounds__(kPlanThreads)
inverse_cdf_plan_kernel(int b, int num_rays, long long valid_rays, int ray_chunk, int P, int max_steps,
                        float fixed_step_size, const int* __restrict__ pts_idx, const float* __restrict__ min_depth,
                        const float* __restrict__ max_depth, const float* __restrict__ noise, float noise_const,
                        const float* __restrict__ probs, const float* __restrict__ steps, int* __restrict__ ray_len,
                        int2* __restrict__ quirk, int* meta /* [max_len, holes, fallback rays] */) {
  const long long ray = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
  int lastv = 0;
  if (ray < valid_rays) {
    const long long H = ray * P;
    const long long batch = ray / num_rays;
    const int rr = (int)(ray - batch * num_rays);
    const int c0 = (rr / ray_chunk) * ray_chunk;
    const long long row0 = batch * num_rays + c0;
    const int* row0_idx = pts_idx + (row0 < valid_rays ? row0 : 0) * P;
    long long nxt = -1;
    if (rr + 1 < min(c0 + ray_chunk, num_rays)) nxt = ray + 1;
    else if (batch + 1 < b) nxt = (batch + 1) * num_rays + c0;
    const int next_idx0 = nxt >= 0 ? pts_idx[(nxt < valid_rays ? nxt : 0) * P] : -1;
    // hit lists are -1-terminated (sorted by the intersection): the first -1 by bisection.  nb: usable bins of this
    // ray (bin 0 is always used); row0_nb: leading valid bins of ray 0 of the block row (its stop test reads THAT ray)
    int nb, row0_nb;
    {
      int lo = 1, hi = P;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (pts_idx[H + mid] == -1) hi = mid; else lo = mid + 1; }
      nb = lo;
      lo = 0; hi = P;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (row0_idx[mid] == -1) hi = mid; else lo = mid + 1; }
      row0_nb = lo;
    }
    const float* noise_row = noise != nullptr ? noise + ray * max_steps : nullptr;
    const float sj = steps[ray];
    const float step_size = fixed_step_size > 0.0f ? fixed_step_size : __fdiv_rn(1.0f, sj);
    const int total_steps = min((int)ceilf(sj), max_steps);
    // pass 1: the reference's left-to-right cumulative sum; non-negative terms => non-decreasing, finite iff the last is
    float cum_last = probs[H];
    bool neg = !(cum_last >= 0.0f);
    for (int j = 1; j < nb; ++j) {
      const float pr = probs[H + j];
      neg |= !(pr >= 0.0f);
      cum_last = __fadd_rn(cum_last, pr);
    }
    const bool ok = !neg && fabsf(cum_last) <= 3.0e38f;
    int n_valid = 0;
    if (ok) {
      // the main loop stops at the first step whose cdf lies beyond the last cumulative sum
      int n_in = total_steps, done = 0;
      if (total_steps > 0 && cdf_at(total_steps - 1, noise_row, noise_const, max_steps, step_size) > cum_last) {
        int lo = 0, hi = total_steps - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cdf_at(mid, noise_row, noise_const, max_steps, step_size) > cum_last) hi = mid; else lo = mid + 1;
        }
        n_in = lo;
        done = 1;
      }
      // pass 2: bin and depth of the last step of the main loop
      int b_prev = -1;
      float z_prev = 0.f;
      if (n_in > 0) {
        const float cl = cdf_at(n_in - 1, noise_row, noise_const, max_steps, step_size);
        float cmin = 0.0f, cmax = probs[H];
        int j = 0;
        while (j < nb - 1 && cl > cmax) {
          cmin = cmax;
          ++j;
          cmax = __fadd_rn(cmax, probs[H + j]);
        }
        if (cl > cmax) j = nb;      // cannot happen for an in-range step; mirrors lower_bound's "none"
        b_prev = j;
        if (j < nb) {
          const float u = __fdiv_rn(__fsub_rn(cl, cmin), __fsub_rn(cmax, cmin));
          const float lo_d = min_depth[H + j];
          z_prev = __fmaf_rn(u, __fsub_rn(max_depth[H + j], lo_d), lo_d);
        }
      }
      const int b_last = done ? nb : (n_in > 0 ? b_prev : 0);
      // step samples and bin ends of the main loop: positions [0, n_in + b_last), all with a valid voxel id
      lastv = n_valid = min(n_in + b_last, max_steps);
      // the reference's trailing loop (sample_gpu.cu:187-200) from the state the main loop leaves
      int curr_bin, sidx = n_in + b_last;
      float curr_max, zl;
      if (done) {
        curr_bin = nb;
        curr_max = max_depth[H + nb - 1];
        zl = b_prev == nb - 1 ? z_prev : min_depth[H + nb - 1];
      } else {
        curr_bin = b_last;
        curr_max = max_depth[H + curr_bin];
        zl = n_in > 0 ? z_prev : min_depth[H];
      }
      while (zl < curr_max) {
        const int oi = curr_bin < P ? pts_idx[H + curr_bin] : next_idx0;
        if (sidx < max_steps && oi != -1) { ++n_valid; lastv = sidx + 1; }
        ++curr_bin;
        ++sidx;
        if (curr_bin >= P || curr_bin >= row0_nb) break;
        curr_max = max_depth[H + curr_bin];
        zl = min_depth[H + curr_bin];
      }
    } else {
      // irregular cumulative sums: count with the reference's serial machine
      CdfState st;
      st.curr_bin = 0; st.s = 0; st.curr_step = 0;
      st.curr_min_depth = min_depth[H]; st.curr_max_depth = max_depth[H];
      st.curr_min_cdf = 0.0f; st.curr_max_cdf = probs[H];
      st.step_size = step_size; st.z_low = st.curr_min_depth; st.total_steps = total_steps;
      st.curr_cdf = 0.0f; st.phase = 0;
      int pos = 0;
      while (pos < max_steps && st.phase != 3) {
        int oi; float od, oz;
        if (cdf_next(st, P, max_steps, H, next_idx0, pts_idx, row0_idx, min_depth, max_depth, probs, noise_row,
                     noise_const, oi, od, oz)) {
          if (oi != -1) { ++n_valid; lastv = pos + 1; }
          ++pos;
        }
      }
      atomicAdd(meta + 2, 1);
    }
    ray_len[ray] = lastv;
    quirk[ray] = make_int2(next_idx0, ok ? row0_nb : -1 - row0_nb);     // negative (-1 - row0_nb) marks a fallback ray
    if (n_valid != lastv) atomicOr(meta + 1, 1);
  }
  const int warp_max = __reduce_max_sync(NSVF_FULL_MASK, lastv);
  if ((threadIdx.x & 31) == 0 && warp_max > 0) atomicMax(meta, warp_max);
}

// Samples of positions [k0, k1) of the live rays -> planes idxT / depthT / distsT [K][ldb].
__global__ void __launch_bounds__(kLazyWarps * 32)
inverse_cdf_block_kernel(long long B, long long ldb, int P, int max_steps, float fixed_step_size, int k0, int k1,
                         const unsigned char* __restrict__ early_stop, const int* __restrict__ ray_len,
                         const int2* __restrict__ quirk, const int* __restrict__ pts_idx,
                         const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                         const float* __restrict__ noise, long long noise_stride, float noise_const,
                         const float* __restrict__ probs, const float* __restrict__ steps, float pad_depth,
                         int* __restrict__ idxT, float* __restrict__ depthT, float* __restrict__ distsT) {
  extern __shared__ __align__(16) float cdf_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile [kLazyBlock][33] x {idx, depth, dist}, then the warps' bin tables
  int* t_idx = reinterpret_cast<int*>(cdf_smem);
  float* t_dep = cdf_smem + kLazyBlock * 33;
  float* t_dst = cdf_smem + 2 * kLazyBlock * 33;
  __shared__ int s_len[32];
  float* base = cdf_smem + 3 * kLazyBlock * 33 + (size_t)warp * 4 * P;
  int* s_idx = reinterpret_cast<int*>(base);
  float *s_min = base + P, *s_max = base + 2 * P, *s_cum = base + 3 * P;
  const int kb = k1 - k0;
  for (long long r0 = (long long)blockIdx.x * 32; r0 < B; r0 += (long long)gridDim.x * 32) {
    __syncthreads();
    if (warp == 0) {
      int l = 0;
      if (r0 + lane < B && (early_stop == nullptr || early_stop[r0 + lane] == 0)) l = min(ray_len[r0 + lane], k1);
      s_len[lane] = l > k0 ? l : 0;          // number of leading positions of this ray that exist, 0 = nothing in the block
    }
    __syncthreads();
    for (int q = 0; q < 32 / kLazyWarps; ++q) {
      const int rl = warp * (32 / kLazyWarps) + q;
      if (s_len[rl] == 0) continue;
      const long long ray = r0 + rl;
      const long long H = ray * P;
      const int2 qk = quirk[ray];
      const float* noise_row = noise != nullptr ? noise + ray * noise_stride : nullptr;
      auto put = [&](int pos, int oi, float od, float oz) {     // ray_sample's clamp / masking, as in the eager path
        if (pos < k0 || pos >= k1) return;
        od = od < 0.f ? 0.f : od;
        if (oi == -1) { od = 0.f; oz = pad_depth; }
        const int a = (pos - k0) * 33 + rl;
        t_idx[a] = oi; t_dep[a] = oz; t_dst[a] = od;
      };
      if (qk.y < 0) {
        // fallback ray: the serial machine from the start, keeping the block's positions (lane 0)
        if (lane == 0) {
          const int* row0_idx = pts_idx;   // unreachable in practice for sliced rays: fallback rays are reported to the host
          CdfState st;
          st.curr_bin = 0; st.s = 0; st.curr_step = 0;
          st.curr_min_depth = min_depth[H]; st.curr_max_depth = max_depth[H];
          st.curr_min_cdf = 0.0f; st.curr_max_cdf = probs[H];
          const float sj = steps[ray];
          st.step_size = fixed_step_size > 0.0f ? fixed_step_size : __fdiv_rn(1.0f, sj);
          st.z_low = st.curr_min_depth; st.total_steps = min((int)ceilf(sj), max_steps);
          st.curr_cdf = 0.0f; st.phase = 0;
          int pos = 0;
          while (pos < min(max_steps, k1) && st.phase != 3) {
            int oi; float od, oz;
            if (cdf_next(st, P, max_steps, H, qk.x, pts_idx, row0_idx, min_depth, max_depth, probs, noise_row,
                         noise_const, oi, od, oz)) {
              put(pos, oi, od, oz);
              ++pos;
            }
          }
        }
        __syncwarp();
        continue;
      }
      const CdfRay r = cdf_ray_setup(H, P, max_steps, fixed_step_size, steps[ray], pts_idx, min_depth, max_depth, probs,
                                     noise_row, noise_const, s_idx, s_min, s_max, s_cum);
      // first step whose position f(i) = i + b_i is >= k0 (f is strictly increasing)
      int i_lo = 0;
      if (lane == 0) {
        int lo = 0, hi = r.n_in;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int bm = bin_of(cdf_at(mid, noise_row, noise_const, max_steps, r.step_size), s_cum, r.nb);
          if (mid + bm >= k0) hi = mid; else lo = mid + 1;
        }
        i_lo = lo;
      }
      i_lo = __shfl_sync(NSVF_FULL_MASK, i_lo, 0);
      // state of step i_lo - 1 (bin and depth), needed for z_low of the first step of the block and for the bin ends
      int pb0 = -1;
      float pz0 = 0.f;
      if (i_lo > 0) {
        const float c = cdf_at(i_lo - 1, noise_row, noise_const, max_steps, r.step_size);
        pb0 = bin_of(c, s_cum, r.nb);
        pz0 = z_at(c, pb0, s_cum, s_min, s_max);
      }
      // (a) step samples of the block, 32 per iteration
      {
        int prev_b = pb0;
        float prev_z = pz0;
        bool past = false;
        for (int i0 = i_lo; i0 < r.n_in && !past; i0 += 32) {
          const int i = i0 + lane;
          const bool act = i < r.n_in;
          float cdf = 0.f, z = 0.f;
          int bb = 0;
          if (act) {
            cdf = cdf_at(i, noise_row, noise_const, max_steps, r.step_size);
            bb = bin_of(cdf, s_cum, r.nb);
            z = z_at(cdf, bb, s_cum, s_min, s_max);
          }
          float pz = __shfl_up_sync(NSVF_FULL_MASK, z, 1);
          int pb = __shfl_up_sync(NSVF_FULL_MASK, bb, 1);
          if (lane == 0) { pz = prev_z; pb = prev_b; }
          if (act) {
            const float zlow = (pb == bb) ? pz : s_min[bb];
            const int pos = i + bb;
            if (pos < max_steps) put(pos, s_idx[bb], __fsub_rn(z, zlow), __fmul_rn(__fadd_rn(z, zlow), 0.5f));
          }
          const unsigned amask = __ballot_sync(NSVF_FULL_MASK, act);
          const int la = 31 - __clz(amask);
          prev_b = __shfl_sync(NSVF_FULL_MASK, bb, la);
          prev_z = __shfl_sync(NSVF_FULL_MASK, z, la);
          past = (i0 + la + prev_b) >= k1;
        }
      }
      // (b) bin-end samples of the block: bin j sits at j + c_j, c_j = first step whose cdf exceeds cum[j]
      {
        bool past = false;
        for (int j0 = (pb0 < 0 ? 0 : pb0); j0 < r.b_last && !past; j0 += 32) {
          const int j = j0 + lane;
          int pos = 0x7fffffff;
          if (j < r.b_last) {
            int lo = 0, hi = r.n_in;              // c_j
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (cdf_at(mid, noise_row, noise_const, max_steps, r.step_size) > s_cum[j]) hi = mid; else lo = mid + 1;
            }
            pos = j + lo;
            if (pos >= k0 && pos < k1 && pos < max_steps) {
              float zlow = s_min[j];
              if (lo > 0) {
                const float c = cdf_at(lo - 1, noise_row, noise_const, max_steps, r.step_size);
                if (bin_of(c, s_cum, r.nb) == j) zlow = z_at(c, j, s_cum, s_min, s_max);
              }
              put(pos, s_idx[j], __fsub_rn(s_max[j], zlow), __fmul_rn(__fadd_rn(s_max[j], zlow), 0.5f));
            }
          }
          past = __any_sync(NSVF_FULL_MASK, j < r.b_last && pos >= k1);
        }
      }
      // (c) trailing samples (lane-uniform replay; every lane computes, lane 0 stores)
      if (r.n_in + r.b_last < k1) {
        cdf_trailing(r, P, noise_row, noise_const, max_steps, qk.y, qk.x, s_idx, s_min, s_max, s_cum, min_depth + H,
                     max_depth + H,
                     [&](int pos, int oi, float od, float oz) { if (lane == 0 && pos < max_steps) put(pos, oi, od, oz); });
      }
      __syncwarp();
    }
    __syncthreads();
    // flush: plane row k of the 32 rays = 128 contiguous bytes
    for (int kk = warp; kk < kb; kk += kLazyWarps) {
      const int k = k0 + kk;
      if (k < s_len[lane]) {
        const long long a = (long long)k * ldb + r0 + lane;
        idxT[a] = t_idx[kk * 33 + lane];
        depthT[a] = t_dep[kk * 33 + lane];
        distsT[a] = t_dst[kk * 33 + lane];
      }
    }
  }
}

// ---- uniform ray sampling -----------------------------------------------------------------------------------------
// Replaces uniform_ray_sampling_kernel, fairnr/clib/src/sample_gpu.cu:15-106.  The reference runs two passes per ray
// IN PLACE in global memory: (1) a three-way merge of voxel entries, voxel exits and march points into the output row,
// (2) a sweep over adjacent pairs of that row that turns them into mid-points / lengths, keeps the pairs lying inside
// a voxel and compacts them to the front — every row is written, re-read and re-written with a stride of
// 4 * max_steps bytes between the threads of a warp.
// Here the two passes are ONE streaming state machine: pass 2 only ever looks at the previous merge point, so each new
// merge point is paired with its predecessor on the spot and the kept sample is emitted directly; the un-compacted
// row never exists.  A lane owns a ray, the warp stages its 32 x TS samples in shared memory and flushes them with
// coalesced stores (TS consecutive slots of 32/TS rays per instruction), and the kernel writes the padding itself.
// What is defined differently from the reference: slots beyond a ray's samples hold
// (idx -1, depth 0, dists 0) instead of stale merge values; reads before the first / past the last element of
// pts_idx yield -1.
struct UniState {
  int s, ucur, umin, umax;        // merge: points produced, march index, next entry, next exit
  int umin2, umax2;               // pair sweep: entries / exits passed (reference phase 2 counters)
  int prev_idx;                   // voxel id attached to the previous merge point
  float prev_depth, base;         // previous merge point, min_depth of the first bin
  bool done;
};

// Produces at most one kept sample per call; false = the ray is finished.
__device__ __forceinline__ bool uni_next(UniState& st, int P, int max_steps, float step_size, long long H,
                                         long long total_hits, const int* __restrict__ pts_idx,
                                         const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                                         const float* __restrict__ noise_row, int& o_idx, float& o_depth,
                                         float& o_dist) {
  while (!st.done) {
    // ---- next merge point (reference :47-81) ----
    if (st.umax == P || st.ucur == max_steps || st.s >= max_steps || __ldg(pts_idx + H + st.umax) == -1) break;
    const float last_min = st.umin < P ? __ldg(min_depth + H + st.umin) : 10000.0f;
    const float last_max = __ldg(max_depth + H + st.umax);
    const float curr = __fmaf_rn(__fadd_rn((float)st.ucur, __ldg(noise_row + st.ucur)), step_size, st.base);
    float d;
    int id;
    if (last_max <= curr && last_max <= last_min) {
      d = last_max; id = __ldg(pts_idx + H + st.umax); st.umax++;
    } else if (curr <= last_min && curr <= last_max) {
      const long long f = H + st.umin - 1;     // umin == 0: the previous ray's last slot (reference quirk)
      d = curr; id = (f >= 0 && f < total_hits) ? __ldg(pts_idx + f) : -1; st.ucur++;
    } else if (last_min <= curr && last_min <= last_max) {
      d = last_min; id = st.umin < P ? __ldg(pts_idx + H + st.umin) : -1; st.umin++;
    } else {
      break;                                    // NaN inputs: the reference would spin forever
    }
    const int pos = st.s++;
    if (pos == 0) { st.prev_depth = d; st.prev_idx = id; continue; }
    // ---- pair (pos - 1, pos): reference phase 2 (:83-100), iteration ucur = pos - 1 ----
    if (id == -1) break;                        // sampled_idx[ucur + 1] == -1
    const float l = st.prev_depth, mid = __fmul_rn(__fadd_rn(l, d), 0.5f), len = __fsub_rn(d, l);
    const int left_idx = st.prev_idx;
    st.prev_depth = d;
    st.prev_idx = id;
    if (st.umin2 < P && mid >= __ldg(min_depth + H + st.umin2) && __ldg(pts_idx + H + st.umin2) > -1) st.umin2++;
    if (st.umax2 < P && mid >= __ldg(max_depth + H + st.umax2) && __ldg(pts_idx + H + st.umax2) > -1) st.umax2++;
    if (st.umax2 == P || __ldg(pts_idx + H + st.umax2) == -1) break;
    if (st.umin2 - 1 == st.umax2 && len > 0.f) {
      o_idx = left_idx; o_depth = mid; o_dist = len;
      return true;
    }
  }
  st.done = true;
  return false;
}

template <int TS>
__global__ void __launch_bounds__(kSampWarps * 32)
uniform_sampling_kernel(long long total_rays, int P, int max_steps, float step_size, const int* __restrict__ pts_idx,
                        const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                        const float* __restrict__ noise, int* __restrict__ sampled_idx,
                        float* __restrict__ sampled_depth, float* __restrict__ sampled_dists,
                        int* __restrict__ max_count) {
  constexpr int LD = TS + 1;
  __shared__ int t_idx[kSampWarps][32 * LD];
  __shared__ float t_depth[kSampWarps][32 * LD];
  __shared__ float t_dist[kSampWarps][32 * LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_hits = total_rays * P;
  const long long n_groups = (total_rays + 31) / 32;
  int my_max = 0;
  for (long long g = (long long)blockIdx.x * kSampWarps + warp; g < n_groups; g += (long long)gridDim.x * kSampWarps) {
    const long long ray = g * 32 + lane;
    const long long H = ray * P;
    UniState st;
    st.s = st.ucur = st.umin = st.umax = st.umin2 = st.umax2 = 0;
    st.prev_idx = -1;
    st.prev_depth = 0.f;
    st.done = ray >= total_rays;
    st.base = st.done ? 0.f : __ldg(min_depth + H);
    const float* noise_row = noise + ray * max_steps;
    const long long out_base = g * 32 * (long long)max_steps;
    const int rows = (int)min((long long)32, total_rays - g * 32);
    int n_valid = 0;
    for (int t0 = 0; t0 < max_steps; t0 += TS) {
      const int tw = min(TS, max_steps - t0);
      int c = 0;
      while (c < tw) {
        int oi; float oz, od;
        if (!uni_next(st, P, max_steps, step_size, H, total_hits, pts_idx, min_depth, max_depth, noise_row, oi, oz, od))
          break;
        t_idx[warp][lane * LD + c] = oi;
        t_depth[warp][lane * LD + c] = oz;
        t_dist[warp][lane * LD + c] = od;
        n_valid += (oi != -1);
        ++c;
      }
      for (; c < tw; ++c) {
        t_idx[warp][lane * LD + c] = -1;
        t_depth[warp][lane * LD + c] = 0.0f;
        t_dist[warp][lane * LD + c] = 0.0f;
      }
      __syncwarp();
      constexpr int RPI = 32 / TS;       // rows per store instruction
      const int cc = lane % TS, rsub = lane / TS;
      for (int r = 0; r < rows; r += RPI) {
        const int rr = r + rsub;
        if (rr < rows && cc < tw) {
          const long long o = out_base + (long long)rr * max_steps + t0 + cc;
          sampled_idx[o] = t_idx[warp][rr * LD + cc];
          sampled_depth[o] = t_depth[warp][rr * LD + cc];
          sampled_dists[o] = t_dist[warp][rr * LD + cc];
        }
      }
      __syncwarp();
      if (__all_sync(NSVF_FULL_MASK, st.done)) {   // every ray of the warp finished: the rest is padding
        const int tnext = t0 + TS;
        for (int r = 0; r < rows && tnext < max_steps; ++r) {
          const long long o = out_base + (long long)r * max_steps + tnext;
          for (int k = lane; k < max_steps - tnext; k += 32) {
            sampled_idx[o + k] = -1;
            sampled_depth[o + k] = 0.0f;
            sampled_dists[o + k] = 0.0f;
          }
        }
        break;
      }
    }
    my_max = max(my_max, n_valid);
  }
  if (max_count != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(NSVF_FULL_MASK, my_max, o));
    if (lane == 0 && my_max > 0) atomicMax(max_count, my_max);
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_inverse_cdf_sampling_ex(nsvf_stream_t stream_, int b, int num_rays, long long valid_rays,
                                         int ray_chunk, int max_hits, int max_steps, float fixed_step_size,
                                         const int* pts_idx, const float* min_depth, const float* max_depth,
                                         const float* uniform_noise, float noise_const, const float* probs,
                                         const float* steps, int* sampled_idx, float* sampled_depth,
                                         float* sampled_dists, int* max_count, int* ray_len, int* holes_flag,
                                         float pad_depth, int flags) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && num_rays >= 0 && max_hits >= 0 && max_steps >= 0, "inverse_cdf_sampling: negative size");
  if (b == 0 || num_rays == 0 || max_steps == 0) return 0;
  NSVF_REQUIRE(max_hits > 0, "inverse_cdf_sampling: max_hits must be > 0 (the reference reads slot 0 of every ray)");
  const long long total_rays = (long long)b * num_rays;
  if (valid_rays < 0 || valid_rays > total_rays) valid_rays = total_rays;
  if (ray_chunk <= 0 || ray_chunk > num_rays) ray_chunk = num_rays;
  if (valid_rays == 0) return 0;
  {  // warp-per-ray kernel whenever its per-warp bin tables fit in shared memory
    const size_t smem = (size_t)kCdfWarps * 6 * max_hits * sizeof(float);
    if (smem <= 160 * 1024 && getenv("NSVF_SAMPLER_LANE") == nullptr) {
      NSVF_CUDA_OK(cudaFuncSetAttribute(inverse_cdf_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          160 * 1024));
      int per_sm = (int)((200 * 1024) / (smem + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 12 ? 12 : per_sm);
      long long want = (valid_rays + kCdfWarps - 1) / kCdfWarps, cap = (long long)num_sms() * per_sm;
      int grid = (int)(want < cap ? want : cap);
      NSVF_TIMED_LAUNCH("inverse_cdf_sampling_kernel", stream,
                        (inverse_cdf_warp_kernel<<<grid, kCdfWarps * 32, smem, stream>>>(
                            b, num_rays, valid_rays, ray_chunk, max_hits, max_steps, fixed_step_size, pts_idx, min_depth,
                            max_depth, uniform_noise, noise_const, probs, steps, sampled_idx, sampled_depth,
                            sampled_dists, max_count, ray_len, holes_flag, pad_depth, flags)));
      return 0;
    }
  }
  const long long groups = (valid_rays + 31) / 32;
  long long want = (groups + kSampWarps - 1) / kSampWarps;
  long long cap = (long long)num_sms() * 8;
  int grid = (int)(want < cap ? want : cap);
  NSVF_TIMED_LAUNCH("inverse_cdf_sampling_kernel", stream, (inverse_cdf_sampling_kernel<16><<<grid, kSampWarps * 32, 0, stream>>>(
      b, num_rays, valid_rays, ray_chunk, max_hits, max_steps, fixed_step_size, pts_idx, min_depth, max_depth,
      uniform_noise, noise_const, probs, steps, sampled_idx, sampled_depth, sampled_dists, max_count, ray_len,
      holes_flag, pad_depth, flags)));
  return 0;
}

extern "C" int nsvf_inverse_cdf_sampling(nsvf_stream_t stream, int b, int num_rays, long long valid_rays,
                                         int ray_chunk, int max_hits, int max_steps, float fixed_step_size,
                                         const int* pts_idx, const float* min_depth, const float* max_depth,
                                         const float* uniform_noise, float noise_const, const float* probs,
                                         const float* steps, int* sampled_idx, float* sampled_depth,
                                         float* sampled_dists, int* max_count) {
  return nsvf_inverse_cdf_sampling_ex(stream, b, num_rays, valid_rays, ray_chunk, max_hits, max_steps, fixed_step_size,
                                      pts_idx, min_depth, max_depth, uniform_noise, noise_const, probs, steps,
                                      sampled_idx, sampled_depth, sampled_dists, max_count, nullptr, nullptr, 0.0f, 1);
}

extern "C" int nsvf_uniform_ray_sampling(nsvf_stream_t stream_, int b, int num_rays, int max_hits, int max_steps,
                                         float step_size, const int* pts_idx, const float* min_depth,
                                         const float* max_depth, const float* uniform_noise, int* sampled_idx,
                                         float* sampled_depth, float* sampled_dists, int* max_count) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && num_rays >= 0 && max_hits >= 0 && max_steps >= 0, "uniform_ray_sampling: negative size");
  if (b == 0 || num_rays == 0 || max_steps == 0) return 0;
  NSVF_REQUIRE(max_hits > 0, "uniform_ray_sampling: max_hits must be > 0");
  const long long total_rays = (long long)b * num_rays;
  const long long groups = (total_rays + 31) / 32;
  long long want = (groups + kSampWarps - 1) / kSampWarps, cap = (long long)num_sms() * 8;
  int grid = (int)(want < cap ? want : cap);
  NSVF_TIMED_LAUNCH("uniform_sampling_kernel", stream,
                    (uniform_sampling_kernel<16><<<grid, kSampWarps * 32, 0, stream>>>(
                        total_rays, max_hits, max_steps, step_size, pts_idx, min_depth, max_depth, uniform_noise,
                        sampled_idx, sampled_depth, sampled_dists, max_count)));
  return 0;
}

extern "C" int nsvf_inverse_cdf_plan(nsvf_stream_t stream_, int b, int num_rays, long long valid_rays, int ray_chunk,
                                     int max_hits, int max_steps, float fixed_step_size, const int* pts_idx,
                                     const float* min_depth, const float* max_depth, const float* uniform_noise,
                                     float noise_const, const float* probs, const float* steps, int* ray_len,
                                     int* quirk, int* meta) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && num_rays >= 0 && max_hits > 0 && max_steps >= 0, "inverse_cdf_plan: bad sizes");
  const long long total_rays = (long long)b * num_rays;
  if (valid_rays < 0 || valid_rays > total_rays) valid_rays = total_rays;
  if (ray_chunk <= 0 || ray_chunk > num_rays) ray_chunk = num_rays;
  if (valid_rays == 0) return 0;
  NSVF_TIMED_LAUNCH("inverse_cdf_plan_kernel", stream,
                    (inverse_cdf_plan_kernel<<<(unsigned)((valid_rays + kPlanThreads - 1) / kPlanThreads), kPlanThreads, 0,
                                               stream>>>(
                        b, num_rays, valid_rays, ray_chunk, max_hits, max_steps, fixed_step_size, pts_idx, min_depth,
                        max_depth, uniform_noise, noise_const, probs, steps, ray_len, reinterpret_cast<int2*>(quirk),
                        meta)));
  return 0;
}

extern "C" long long nsvf_march_plane_stride(long long B);

extern "C" int nsvf_inverse_cdf_block(nsvf_stream_t stream_, long long B, int max_hits, int max_steps,
                                      float fixed_step_size, int k_begin, int k_end, const unsigned char* early_stop,
                                      const int* ray_len, const int* quirk, const int* pts_idx, const float* min_depth,
                                      const float* max_depth, const float* uniform_noise, long long noise_row_stride,
                                      float noise_const, const float* probs, const float* steps, float pad_depth,
                                      int* idxT, float* depthT, float* distsT) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && max_hits > 0 && max_steps >= 0 && k_begin >= 0 && k_begin <= k_end, "inverse_cdf_block: bad sizes");
  if (B == 0 || k_end == k_begin) return 0;
  const size_t smem = ((size_t)3 * kLazyBlock * 33 + (size_t)kLazyWarps * 4 * max_hits) * sizeof(float);
  NSVF_REQUIRE(smem <= 200 * 1024, "inverse_cdf_block: max_hits=%d too large", max_hits);
  NSVF_CUDA_OK(cudaFuncSetAttribute(inverse_cdf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  const long long ldb = nsvf_march_plane_stride(B);
  for (int k0 = k_begin; k0 < k_end; k0 += kLazyBlock) {
    const int k1 = k0 + kLazyBlock < k_end ? k0 + kLazyBlock : k_end;
    long long want = (B + 31) / 32, cap = (long long)num_sms() * per_sm;
    NSVF_TIMED_LAUNCH("inverse_cdf_block_kernel", stream,
                      (inverse_cdf_block_kernel<<<(int)(want < cap ? want : cap), kLazyWarps * 32, smem, stream>>>(
                          B, ldb, max_hits, max_steps, fixed_step_size, k0, k1, early_stop, ray_len,
                          reinterpret_cast<const int2*>(quirk), pts_idx, min_depth, max_depth, uniform_noise,
                          noise_row_stride, noise_const, probs, steps, pad_depth, idxT, depthT, distsT)));
  }
  return 0;
}

extern "C" size_t nsvf_inverse_cdf_stream_state_bytes(long long B) { return (size_t)(B > 0 ? B : 0) * sizeof(CdfParked); }

extern "C" int nsvf_inverse_cdf_stream(nsvf_stream_t stream_, long long B, int max_hits, int max_steps,
                                       float fixed_step_size, int k_begin, int k_end, const unsigned char* early_stop,
                                       const int* ray_len, const int* quirk, const int* pts_idx, const float* min_depth,
                                       const float* max_depth, const float* uniform_noise, long long noise_row_stride,
                                       float noise_const, const float* probs, const float* steps, float pad_depth,
                                       void* state, int* idxT, float* depthT, float* distsT) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && max_hits > 0 && max_steps >= 0 && k_begin >= 0 && k_begin <= k_end, "inverse_cdf_stream: bad sizes");
  NSVF_REQUIRE(state != nullptr && ((uintptr_t)state & 15) == 0, "inverse_cdf_stream: state must be a 16-byte aligned buffer");
  if (B == 0 || k_end == k_begin) return 0;
  NSVF_TIMED_LAUNCH("inverse_cdf_stream_kernel", stream,
                    (inverse_cdf_stream_kernel<<<(unsigned)((B + kStreamThreads - 1) / kStreamThreads), kStreamThreads, 0,
                                                 stream>>>(
                        B, nsvf_march_plane_stride(B), max_hits, max_steps, fixed_step_size, k_begin, k_end, early_stop,
                        ray_len, reinterpret_cast<const int2*>(quirk), pts_idx, min_depth, max_depth, uniform_noise,
                        noise_row_stride, noise_const, probs, steps, pad_depth, reinterpret_cast<CdfParked*>(state), idxT,
                        depthT, distsT)));
  return 0;
}
