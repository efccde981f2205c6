// Device-side ray-marching plan for the renderer (sm_100a): chunk schedule, sample compaction, free-energy
// epilogue with early termination — the loop of VolumeRenderer.forward_chunk / forward_once,
// fairnr/modules/renderer.py:77-191, without its per-column host syncs and dense [B,K] scatter traffic.
//
// The reference walks the sample columns, sums `hits[:, i]` on the host for every column (one sync each), evaluates
// the field on the boolean-compacted samples of a column window whenever the running count would exceed chunk_size,
// masked_scatter()s sigma / texture back into zero-filled [B,K] tensors, and after every window re-derives which
// rays have accumulated enough free energy to stop.  Here the same schedule — the identical windows, hence the
// identical set of samples reaching the field — is kept on the device:
//
//   * samples live in TRIMMED rows: ray r owns slots [0, len_r) of its row (row stride ldk), nothing beyond len_r is
//     ever read or written, so traffic is proportional to the samples that exist, not to B x K;
//   * per-column counts of live samples are maintained incrementally: a histogram of the row lengths at the start,
//     and a (-1 at the window end, +1 at len_r) difference pair for every ray that stops; the LAST CTA of the
//     epilogue kernel integrates the differences and picks the next window (same rule as renderer.py:157-158), so
//     one launch does scatter + early-stop + scheduling; the host reads 4 ints per window (or, without early
//     termination, the whole window list once);
//   * compaction is a single-pass kernel: per-ray counts are pure arithmetic on (len_r, window, alive), tile offsets
//     come from a decoupled look-back scan (ticketed tiles, epoch-tagged 64-bit states: no reset between launches).
//
// Precondition of this path: the valid samples of every ray form a prefix of its row (true for both samplers; the
// sampler reports rows where it is not, nsvf_march_ray_lengths checks foreign inputs) — otherwise the caller uses
// the general compaction kernels of compact.cu.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

// plan layout (int32 words, device memory, zero-initialised by the caller once per forward_chunk):
//   [0..15]            header: 0 start, 1 end, 2 count, 3 done, 4 holes, 5 n_windows, 6 ticket, 7 ctas_done,
//                              8 total_samples
//   [16 .. 16+K)       counts[k]  = live rays with a valid sample in column k
//   [16+K .. 16+2K+1)  diff[k]    = pending difference array for counts (prefix-summed from the window start)
//   then (8-byte aligned) tile_state[n_tiles] u64 for the look-back scan
constexpr int kHdr = 16;
constexpr int H_START = 0, H_END = 1, H_COUNT = 2, H_DONE = 3, H_HOLES = 4, H_NWIN = 5, H_TICKET = 6, H_CTAS = 7,
              H_TOTAL = 8;
constexpr int kTile = 256;        // rays per compaction tile
constexpr int kNarrow = 12;       // windows up to this many columns: one thread per ray; wider: one warp per ray

__host__ __device__ inline long long plan_tiles(long long B) { return (B + kTile - 1) / kTile; }
__host__ __device__ inline size_t plan_words_before_tiles(int K) {
  size_t w = (size_t)kHdr + (size_t)K + (size_t)K + 1;
  return (w + 1) & ~(size_t)1;   // 8-byte alignment for the u64 tile states
}

__device__ __forceinline__ int warp_incl_sum_i(int x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(NSVF_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  return x;
}

// Executed by ONE warp.  Integrates the pending differences into counts[from..K) and selects the window that starts
// at `from`: the longest [from, end) whose live-sample count stays <= chunk_size (at least one column), exactly the
// flush rule of renderer.py:157-158.  Returns (end, count) in all lanes.
__device__ __forceinline__ void select_window(int* counts, int* diff, int K, int from, int chunk_size, bool integrate,
                                              int& end, int& count) {
  const int lane = threadIdx.x & 31;
  if (integrate) {
    int run = 0;
    for (int k0 = from; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      const int d = k < K ? diff[k] : 0;
      const int incl = warp_incl_sum_i(d, lane);
      if (k < K) {
        counts[k] += run + incl;
        diff[k] = 0;
      }
      run += __shfl_sync(NSVF_FULL_MASK, incl, 31);
    }
    __syncwarp();
  }
  int total = 0;
  end = K;
  count = 0;
  bool found = false;
  for (int k0 = from; k0 < K && !found; k0 += 32) {
    const int k = k0 + lane;
    const int c = k < K ? counts[k] : 0;
    const int incl = warp_incl_sum_i(c, lane) + total;
    const bool over = (k < K) && (k > from) && (incl > chunk_size);
    const unsigned m = __ballot_sync(NSVF_FULL_MASK, over);
    if (m) {
      const int first = __ffs(m) - 1;
      end = k0 + first;
      const int prev = __shfl_sync(NSVF_FULL_MASK, incl, first > 0 ? first - 1 : 0);
      count = first > 0 ? prev : total;
      found = true;
    } else {
      total = __shfl_sync(NSVF_FULL_MASK, incl, 31);
    }
  }
  if (!found) count = total;
}

// One warp: publish the next window (or, with all_windows, every window of the call) to the plan header and to the
// host-visible mirror.  host_info (pinned, device-mapped) layout: [0..8] header copy, then triplets (start,end,count).
__device__ void publish_schedule(int* plan, int K, int chunk_size, int from, bool integrate, bool all_windows,
                                 volatile int* host_info, int host_capacity) {
  const int lane = threadIdx.x & 31;
  int* counts = plan + kHdr;
  int* diff = plan + kHdr + K;
  int end, count;
  if (!all_windows) {
    select_window(counts, diff, K, from, chunk_size, integrate, end, count);
    const int done = (from >= K || count == 0) ? 1 : 0;
    if (lane == 0) {
      plan[H_START] = from; plan[H_END] = end; plan[H_COUNT] = count; plan[H_DONE] = done;
      if (host_info != nullptr) {
        host_info[H_START] = from; host_info[H_END] = end; host_info[H_COUNT] = count; host_info[H_DONE] = done;
        host_info[H_HOLES] = plan[H_HOLES]; host_info[H_TOTAL] = plan[H_TOTAL];
      }
    }
  } else {
    int n = 0, s = from;
    bool first = true;
    while (s < K) {
      select_window(counts, diff, K, s, chunk_size, integrate && first, end, count);
      first = false;
      if (count == 0) break;
      if (lane == 0 && host_info != nullptr && kHdr + 3 * n + 2 < host_capacity) {
        host_info[kHdr + 3 * n + 0] = s; host_info[kHdr + 3 * n + 1] = end; host_info[kHdr + 3 * n + 2] = count;
      }
      ++n;
      s = end;
    }
    if (lane == 0) {
      plan[H_NWIN] = n; plan[H_DONE] = 1;
      if (host_info != nullptr) {
        host_info[H_NWIN] = n; host_info[H_DONE] = 1; host_info[H_HOLES] = plan[H_HOLES];
        host_info[H_TOTAL] = plan[H_TOTAL];
      }
    }
  }
  __threadfence_system();
}

// "last CTA" election: returns true in every thread of the CTA that finishes last (and resets the counter).
__device__ __forceinline__ bool last_cta(int* plan) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(plan + H_CTAS, 1);
    s_last = (prev == (int)gridDim.x - 1);
    if (s_last) plan[H_CTAS] = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ---- row lengths of foreign (padded) sample tensors ------------------------------------------------------------
// lens[r] = 1 + index of the last valid (idx != -1) slot; flags a row whose valid slots are not a prefix.
__global__ void __launch_bounds__(256)
march_ray_lengths_kernel(long long B, int K, long long ldk, const int* __restrict__ idx, int* __restrict__ lens,
                         int* __restrict__ plan) {
  const int lane = threadIdx.x & 31;
  int holes = 0;
  for (long long ray = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); ray < B; ray += (long long)gridDim.x * 8) {
    int last = 0, nvalid = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      const bool ok = k < K && idx[ray * ldk + k] != -1;
      const unsigned m = __ballot_sync(NSVF_FULL_MASK, ok);
      if (m) { last = k0 + 32 - __clz(m); nvalid += __popc(m); }
    }
    if (lane == 0) lens[ray] = last;
    holes |= (nvalid != last);
  }
  if (holes && lane == 0) atomicOr(plan + H_HOLES, 1);
}

// ---- begin: histogram of the row lengths -> counts, first window / all windows ---------------------------------
__global__ void __launch_bounds__(256)
march_begin_kernel(long long B, int K, int chunk_size, const int* __restrict__ lens,
                   const unsigned char* __restrict__ early_stop, int all_windows, int* __restrict__ plan,
                   volatile int* host_info, int host_capacity) {
  extern __shared__ int s_hist[];   // K + 1 bins
  for (int k = threadIdx.x; k <= K; k += blockDim.x) s_hist[k] = 0;
  __syncthreads();
  int* diff = plan + kHdr + K;
  int alive = 0;
  long long samples = 0;
  for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < B;
       ray += (long long)gridDim.x * blockDim.x) {
    if (early_stop != nullptr && early_stop[ray]) continue;
    int l = lens[ray];
    l = l < 0 ? 0 : (l > K ? K : l);
    if (l > 0) { ++alive; samples += l; atomicAdd(s_hist + l, 1); }
  }
  // counts[k] = #{len > k} = prefix sum of (+alive at 0, -1 at len)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    alive += __shfl_xor_sync(NSVF_FULL_MASK, alive, o);
    samples += __shfl_xor_sync(NSVF_FULL_MASK, samples, o);
  }
  if ((threadIdx.x & 31) == 0 && alive) {
    atomicAdd(diff, alive);
    atomicAdd(plan + H_TOTAL, (int)samples);
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= K; k += blockDim.x)
    if (s_hist[k]) atomicSub(diff + k, s_hist[k]);
  if (last_cta(plan) && threadIdx.x < 32)
    publish_schedule(plan, K, chunk_size, 0, true, all_windows != 0, host_info, host_capacity);
}

// ---- single-pass compaction of one column window ----------------------------------------------------------------
__device__ __forceinline__ int window_samples(int len, int start, int end) {
  const int hi = len < end ? len : end;
  return hi > start ? hi - start : 0;
}

__global__ void __launch_bounds__(kTile)
march_compact_kernel(long long B, int K, long long ldk, int start, int end, const int* __restrict__ lens,
                     const unsigned char* __restrict__ early_stop, const int* __restrict__ s_idx,
                     const float* __restrict__ s_depth, const float* __restrict__ s_dists,
                     const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                     int* __restrict__ out_vox, float* __restrict__ out_xyz, float* __restrict__ out_dir,
                     float* __restrict__ out_dists, int* __restrict__ ray_off, int* __restrict__ plan,
                     unsigned long long* __restrict__ tile_state, unsigned epoch, unsigned ticket_base) {
  __shared__ unsigned sh_tile, sh_base;
  __shared__ int sh_warp[kTile / 32];
  __shared__ int sh_off[kTile], sh_n[kTile];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sh_tile = atomicAdd(reinterpret_cast<unsigned*>(plan + H_TICKET), 1u) - ticket_base;
  __syncthreads();
  const unsigned tile = sh_tile;
  const long long ray = (long long)tile * kTile + tid;
  int n = 0;
  if (ray < B && (early_stop == nullptr || early_stop[ray] == 0)) n = window_samples(lens[ray], start, end);
  // block-wide exclusive scan of n
  const int incl = warp_incl_sum_i(n, lane);
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kTile / 32 ? sh_warp[lane] : 0;
    const int wi = warp_incl_sum_i(w, lane);
    if (lane < kTile / 32) sh_warp[lane] = wi - w;          // exclusive warp bases
    if (lane == kTile / 32 - 1) sh_off[0] = wi;             // stash the tile total
  }
  __syncthreads();
  const int tile_total = sh_off[0];
  const int excl = incl - n + sh_warp[warp];
  __syncthreads();
  // decoupled look-back: state = (epoch*4 + flag) << 32 | value, flag 1 = aggregate, 2 = inclusive prefix
  if (tid == 0) {
    const unsigned long long tag = (unsigned long long)epoch << 34;
    unsigned base = 0;
    if (tile == 0) {
      atomicExch(tile_state + tile, tag | (2ull << 32) | (unsigned)tile_total);
    } else {
      atomicExch(tile_state + tile, tag | (1ull << 32) | (unsigned)tile_total);
      long long p = (long long)tile - 1;
      while (true) {
        const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(tile_state + p);
        if ((s >> 34) != epoch) continue;                   // not published in this launch yet
        base += (unsigned)(s & 0xffffffffull);
        if (((s >> 32) & 3ull) == 2ull) break;
        --p;
      }
      atomicExch(tile_state + tile, tag | (2ull << 32) | (unsigned)(base + tile_total));
    }
    sh_base = base;
  }
  __syncthreads();
  const int off = (int)sh_base + excl;
  if (ray < B) ray_off[ray] = off;
  if (ray == B - 1) ray_off[B] = off + n;
  sh_off[tid] = off;
  sh_n[tid] = n;
  __syncthreads();

  if (end - start <= kNarrow) {
    if (n > 0) {
      const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
      const float dx = ray_dir[ray * 3 + 0], dy = ray_dir[ray * 3 + 1], dz = ray_dir[ray * 3 + 2];
      const long long row = ray * ldk + start;
      for (int t = 0; t < n; ++t) {
        const long long o = (long long)off + t;
        const float d = s_depth[row + t];
        out_vox[o] = s_idx[row + t];
        // ray(): ray_start + ray_dir * depth — a separate multiply and add in the reference (no FMA)
        out_xyz[o * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, d));
        out_xyz[o * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, d));
        out_xyz[o * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, d));
        if (out_dir != nullptr) { out_dir[o * 3 + 0] = dx; out_dir[o * 3 + 1] = dy; out_dir[o * 3 + 2] = dz; }
        if (out_dists != nullptr) out_dists[o] = s_dists[row + t];
      }
    }
  } else {
    // wide window: a warp walks its 32 rays one after the other, lanes along the samples (coalesced both ways)
    for (int j = 0; j < 32; ++j) {
      const int lt = warp * 32 + j;
      const int nn = sh_n[lt];
      if (nn == 0) continue;
      const long long r = (long long)tile * kTile + lt;
      const long long o0 = sh_off[lt];
      const float ox = ray_start[r * 3 + 0], oy = ray_start[r * 3 + 1], oz = ray_start[r * 3 + 2];
      const float dx = ray_dir[r * 3 + 0], dy = ray_dir[r * 3 + 1], dz = ray_dir[r * 3 + 2];
      const long long row = r * ldk + start;
      for (int t = lane; t < nn; t += 32) {
        const long long o = o0 + t;
        const float d = s_depth[row + t];
        out_vox[o] = s_idx[row + t];
        out_xyz[o * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, d));
        out_xyz[o * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, d));
        out_xyz[o * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, d));
        if (out_dir != nullptr) { out_dir[o * 3 + 0] = dx; out_dir[o * 3 + 1] = dy; out_dir[o * 3 + 2] = dz; }
        if (out_dists != nullptr) out_dists[o] = s_dists[row + t];
      }
    }
  }
}

// ---- epilogue of one window: free energy, scatter into the trimmed rows, early termination, next window ---------
// free_energy = relu(noise + sigma) * dists * 7   (renderer.py:117-121; op order kept)
__device__ __forceinline__ float free_energy(float sigma, float noise, float dist) {
  const float a = __fadd_rn(noise, sigma);
  return __fmul_rn(__fmul_rn(a > 0.f ? a : 0.f, dist), 7.0f);
}

__global__ void __launch_bounds__(kTile)
march_epilogue_kernel(long long B, int K, int start, int end, const int* __restrict__ ray_off,
                      const int* __restrict__ lens, unsigned char* __restrict__ early_stop,
                      float* __restrict__ acc_fe, int* __restrict__ eval_len, const float* __restrict__ sigma,
                      const float* __restrict__ noise, const float* __restrict__ dists,
                      const float* __restrict__ texture, float tolerance, float* __restrict__ fe_rows,
                      float* __restrict__ tex_rows, int chunk_size, int schedule_next, int* __restrict__ plan,
                      volatile int* host_info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* diff = plan + kHdr + K;
  if (end - start <= kNarrow) {
    for (long long ray = (long long)blockIdx.x * kTile + tid; ray < B; ray += (long long)gridDim.x * kTile) {
      const int o0 = ray_off[ray], n = ray_off[ray + 1] - o0;
      if (n == 0) continue;
      const long long row = ray * K + start;
      float sum = 0.f;
      for (int t = 0; t < n; ++t) {
        const long long o = (long long)o0 + t;
        if (sigma != nullptr) {
          const float fe = free_energy(sigma[o], noise != nullptr ? noise[o] : 0.f, dists[o]);
          fe_rows[row + t] = fe;
          sum += fe;
        }
        if (texture != nullptr) {
          tex_rows[(row + t) * 3 + 0] = texture[o * 3 + 0];
          tex_rows[(row + t) * 3 + 1] = texture[o * 3 + 1];
          tex_rows[(row + t) * 3 + 2] = texture[o * 3 + 2];
        }
      }
      eval_len[ray] = start + n;
      if (tolerance > 0.f && sigma != nullptr) {
        const float acc = acc_fe[ray] + sum;
        acc_fe[ray] = acc;
        if (acc > tolerance) {
          early_stop[ray] = 1;
          const int len = lens[ray];
          if (len > end) { atomicSub(diff + end, 1); atomicAdd(diff + len, 1); }
        }
      }
    }
  } else {
    for (long long ray = (long long)blockIdx.x * (kTile / 32) + warp; ray < B;
         ray += (long long)gridDim.x * (kTile / 32)) {
      const int o0 = ray_off[ray], n = ray_off[ray + 1] - o0;
      if (n == 0) continue;
      const long long row = ray * K + start;
      float sum = 0.f;
      for (int t = lane; t < n; t += 32) {
        const long long o = (long long)o0 + t;
        if (sigma != nullptr) {
          const float fe = free_energy(sigma[o], noise != nullptr ? noise[o] : 0.f, dists[o]);
          fe_rows[row + t] = fe;
          sum += fe;
        }
        if (texture != nullptr) {
          tex_rows[(row + t) * 3 + 0] = texture[o * 3 + 0];
          tex_rows[(row + t) * 3 + 1] = texture[o * 3 + 1];
          tex_rows[(row + t) * 3 + 2] = texture[o * 3 + 2];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(NSVF_FULL_MASK, sum, o);
      if (lane == 0) {
        eval_len[ray] = start + n;
        if (tolerance > 0.f && sigma != nullptr) {
          const float acc = acc_fe[ray] + sum;
          acc_fe[ray] = acc;
          if (acc > tolerance) {
            early_stop[ray] = 1;
            const int len = lens[ray];
            if (len > end) { atomicSub(diff + end, 1); atomicAdd(diff + len, 1); }
          }
        }
      }
    }
  }
  if (schedule_next && last_cta(plan) && tid < 32)
    publish_schedule(plan, K, chunk_size, end, true, false, host_info, 0);
}

// backward of the epilogue for one window: gradients of the trimmed rows back to the compacted field outputs
//   d sigma = (g_fe * 7) * dists * [noise + sigma > 0];  d texture = g_tex
__global__ void __launch_bounds__(256)
march_epilogue_bwd_kernel(long long B, int K, int start, const int* __restrict__ ray_off,
                          const float* __restrict__ g_fe_rows, const float* __restrict__ g_tex_rows,
                          const float* __restrict__ sigma, const float* __restrict__ noise,
                          const float* __restrict__ dists, float* __restrict__ g_sigma, float* __restrict__ g_texture,
                          int wide) {
  const int tid = threadIdx.x, lane = tid & 31;
  const long long stride = wide ? (long long)gridDim.x * 8 : (long long)gridDim.x * 256;
  for (long long ray = wide ? (long long)blockIdx.x * 8 + (tid >> 5) : (long long)blockIdx.x * 256 + tid; ray < B;
       ray += stride) {
    const int o0 = ray_off[ray], n = ray_off[ray + 1] - o0;
    const long long row = ray * K + start;
    for (int t = wide ? lane : 0; t < n; t += wide ? 32 : 1) {
      const long long o = (long long)o0 + t;
      if (g_sigma != nullptr) {
        const float a = __fadd_rn(noise != nullptr ? noise[o] : 0.f, sigma[o]);
        g_sigma[o] = a > 0.f ? __fmul_rn(__fmul_rn(g_fe_rows[row + t], 7.0f), dists[o]) : 0.f;
      }
      if (g_texture != nullptr) {
        g_texture[o * 3 + 0] = g_tex_rows[(row + t) * 3 + 0];
        g_texture[o * 3 + 1] = g_tex_rows[(row + t) * 3 + 1];
        g_texture[o * 3 + 2] = g_tex_rows[(row + t) * 3 + 2];
      }
    }
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" size_t nsvf_march_plan_bytes(long long B, int K) {
  if (B < 0 || K < 0) return 0;
  return plan_words_before_tiles(K) * sizeof(int) + (size_t)plan_tiles(B) * sizeof(unsigned long long);
}

extern "C" int nsvf_march_ray_lengths(nsvf_stream_t stream_, long long B, int K, long long ldk,
                                      const int* sampled_idx, int* lens, void* plan) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && ldk >= K, "march_ray_lengths: bad sizes");
  if (B == 0) return 0;
  long long want = (B + 7) / 8, cap = (long long)num_sms() * 16;
  march_ray_lengths_kernel<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(B, K, ldk, sampled_idx, lens,
                                                                               (int*)plan);
  NSVF_LAUNCH_OK("march_ray_lengths_kernel");
  return 0;
}

extern "C" int nsvf_march_begin(nsvf_stream_t stream_, long long B, int K, int chunk_size, const int* lens,
                                const unsigned char* early_stop, int all_windows, void* plan, int* host_info,
                                int host_capacity_ints) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && chunk_size > 0, "march_begin: bad sizes");
  NSVF_REQUIRE(host_info == nullptr || host_capacity_ints >= kHdr, "march_begin: host_info too small");
  const size_t smem = sizeof(int) * ((size_t)K + 1);
  NSVF_REQUIRE(smem <= 160 * 1024, "march_begin: K=%d too large", K);
  if (smem > 48 * 1024)
    NSVF_CUDA_OK(cudaFuncSetAttribute(march_begin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long want = (B + 255) / 256, cap = (long long)num_sms() * 4;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  march_begin_kernel<<<grid, 256, smem, stream>>>(B, K, chunk_size, lens, early_stop, all_windows, (int*)plan,
                                                  host_info, host_capacity_ints);
  NSVF_LAUNCH_OK("march_begin_kernel");
  return 0;
}

extern "C" int nsvf_march_compact(nsvf_stream_t stream_, long long B, int K, long long ldk, int start, int end,
                                  const int* lens, const unsigned char* early_stop, const int* sampled_idx,
                                  const float* sampled_depth, const float* sampled_dists, const float* ray_start,
                                  const float* ray_dir, int* out_vox, float* out_xyz, float* out_dir,
                                  float* out_dists, int* ray_off, void* plan, int launch_no) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && ldk >= K && start >= 0 && start <= end && end <= K && launch_no >= 0,
               "march_compact: bad sizes");
  if (B == 0) return 0;
  const long long tiles = plan_tiles(B);
  NSVF_REQUIRE(tiles * ((long long)launch_no + 1) < 0xffffffffll, "march_compact: ticket counter would overflow");
  unsigned long long* tile_state =
      reinterpret_cast<unsigned long long*>((int*)plan + plan_words_before_tiles(K));
  NSVF_TIMED_LAUNCH("march_compact_kernel", stream,
                    (march_compact_kernel<<<(unsigned)tiles, kTile, 0, stream>>>(
                        B, K, ldk, start, end, lens, early_stop, sampled_idx, sampled_depth, sampled_dists, ray_start,
                        ray_dir, out_vox, out_xyz, out_dir, out_dists, ray_off, (int*)plan, tile_state,
                        (unsigned)launch_no + 1u, (unsigned)(tiles * launch_no))));
  return 0;
}

extern "C" int nsvf_march_epilogue(nsvf_stream_t stream_, long long B, int K, int start, int end, const int* ray_off,
                                   const int* lens, unsigned char* early_stop, float* acc_free_energy, int* eval_len,
                                   const float* sigma, const float* noise, const float* dists, const float* texture,
                                   float tolerance, float* free_energy_rows, float* texture_rows, int chunk_size,
                                   int schedule_next, void* plan, int* host_info) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && start >= 0 && start <= end && end <= K, "march_epilogue: bad sizes");
  NSVF_REQUIRE(sigma == nullptr || (dists != nullptr && free_energy_rows != nullptr), "march_epilogue: sigma needs dists");
  NSVF_REQUIRE(texture == nullptr || texture_rows != nullptr, "march_epilogue: texture needs texture_rows");
  if (B == 0) return 0;
  const bool narrow = end - start <= kNarrow;
  long long want = narrow ? (B + kTile - 1) / kTile : (B + kTile / 32 - 1) / (kTile / 32);
  long long cap = (long long)num_sms() * 8;
  NSVF_TIMED_LAUNCH("march_epilogue_kernel", stream,
                    (march_epilogue_kernel<<<(int)(want < cap ? want : cap), kTile, 0, stream>>>(
                        B, K, start, end, ray_off, lens, early_stop, acc_free_energy, eval_len, sigma, noise, dists,
                        texture, tolerance, free_energy_rows, texture_rows, chunk_size, schedule_next, (int*)plan,
                        host_info)));
  return 0;
}

extern "C" int nsvf_march_epilogue_bwd(nsvf_stream_t stream_, long long B, int K, int start, int end,
                                       const int* ray_off, const float* grad_free_energy_rows,
                                       const float* grad_texture_rows, const float* sigma, const float* noise,
                                       const float* dists, float* grad_sigma, float* grad_texture) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && start >= 0 && start <= end && end <= K, "march_epilogue_bwd: bad sizes");
  if (B == 0) return 0;
  const int wide = end - start > kNarrow;
  long long want = wide ? (B + 7) / 8 : (B + 255) / 256, cap = (long long)num_sms() * 8;
  march_epilogue_bwd_kernel<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(
      B, K, start, ray_off, grad_free_energy_rows, grad_texture_rows, sigma, noise, dists, grad_sigma, grad_texture,
      wide);
  NSVF_LAUNCH_OK("march_epilogue_bwd_kernel");
  return 0;
}
