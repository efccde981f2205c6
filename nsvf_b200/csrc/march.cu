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
//   * the call's samples are transposed ONCE into slot-major planes [K][B] (march_transpose_kernel, a tiled smem
//     transpose that only touches the lens[r] valid slots of each row).  A window is a range of sample columns, so with
//     one thread per ray every access of the window loop — compaction reads, free-energy / texture writes, the
//     compositing scan and its backward — is coalesced across the warp, and nothing beyond a ray's last sample is
//     ever read or written: traffic is proportional to the samples that exist, not to B x K;
//   * per-column counts of live samples are maintained incrementally: a histogram of the row lengths at the start,
//     and a (-1 at the window end, +1 at len_r) difference pair for every ray that stops; the LAST CTA of the
//     epilogue kernel integrates the differences and picks the next window (same rule as renderer.py:157-158), so
//     one launch does scatter + early-stop + scheduling; the host reads 4 ints per window (or, without early
//     termination, the whole window list once);
//   * compaction: per-ray counts are pure arithmetic on (len_r, window, alive).  A first tiny kernel leaves the sample
//     count of every 256-ray tile in the plan, the second one sums the counts of the tiles before its own (at most a
//     few thousand L2-resident ints) and writes its samples — two short launches with no dependence between CTAs.  The
//     first generation did it in one pass with a decoupled look-back over ticketed tiles: 30 us per window on 2048
//     tiles, most of it the ticket counter and the prefix wave travelling down the chain.
//
// Precondition of this path: the valid samples of every ray form a prefix of its row (true for both samplers; the
// sampler reports rows where it is not, nsvf_march_ray_lengths checks foreign inputs) — otherwise the caller uses
// the general compaction kernels of compact.cu.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

// plan layout (int32 words, device memory, zero-initialised by the caller once per forward_chunk):
//   [0..15]            header: 0 start, 1 end, 2 count, 3 done, 4 holes, 5 n_windows, 6 carry (sum of the pending differences before `start`), 7 ctas_done,
//                              8 total_samples
//   [16 .. 16+K)       counts[k]  = live rays with a valid sample in column k
//   [16+K .. 16+2K+1)  diff[k]    = pending difference array for counts (prefix-summed from the window start)
//   then (8-byte aligned) tile_total[n_tiles] i32: samples of every 256-ray tile in the window being compacted
//   (8 bytes reserved per tile)
constexpr int kHdr = 16;
constexpr int H_START = 0, H_END = 1, H_COUNT = 2, H_DONE = 3, H_HOLES = 4, H_NWIN = 5, H_CARRY = 6, H_CTAS = 7,
              H_TOTAL = 8;
constexpr int kTile = 256;        // rays per compaction tile

// Row stride (in elements) of the slot-major planes: B rounded up to 32 rays, and never a multiple of 1024 — with
// B = 2^19 rays (chunk_size 512) consecutive slots of a ray would lie exactly 2 MiB apart and every access of the
// transpose would fall into the same memory channel.
__host__ __device__ inline long long plane_stride(long long B) {
  long long s = (B + 31) / 32 * 32;
  if (s % 1024 == 0) s += 32;
  return s;
}
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
    // One window at a time (early termination): the pending differences are NOT integrated over all K columns per
    // window (22 dependent rounds of global loads and stores by a single warp at K = 700: half of the epilogue
    // kernel's time).  Windows only move forward and every new difference lands at or behind the current window's
    // end, so it is enough to carry the running sum of the differences in front of `from` in the header and to add
    // the differences of the few columns the window selection actually looks at; counts[] / diff[] stay read-only.
    int run = plan[H_CARRY], total = 0;
    bool found = false;
    end = K;
    count = 0;
    for (int k0 = from; k0 < K && !found; k0 += 32) {
      const int k = k0 + lane;
      const int d = k < K ? diff[k] : 0;
      const int incl_d = warp_incl_sum_i(d, lane);
      const int c = k < K ? counts[k] + run + incl_d : 0;       // live samples in column k
      const int incl = warp_incl_sum_i(c, lane) + total;
      const bool over = (k < K) && (k > from) && (incl > chunk_size);
      const unsigned m = __ballot_sync(NSVF_FULL_MASK, over);
      if (m) {
        const int first = __ffs(m) - 1;
        end = k0 + first;
        const int prev = __shfl_sync(NSVF_FULL_MASK, incl, first > 0 ? first - 1 : 0);
        const int prev_d = __shfl_sync(NSVF_FULL_MASK, incl_d, first > 0 ? first - 1 : 0);
        count = first > 0 ? prev : total;
        run += first > 0 ? prev_d : 0;                             // differences of the columns [from, end)
        found = true;
      } else {
        total = __shfl_sync(NSVF_FULL_MASK, incl, 31);
        run += __shfl_sync(NSVF_FULL_MASK, incl_d, 31);
      }
    }
    if (!found) count = total;
    (void)integrate;
    const int done = (from >= K || count == 0) ? 1 : 0;
    if (lane == 0) {
      plan[H_CARRY] = run;
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

// ---- row-major trimmed rows -> slot-major planes ------------------------------------------------------------------
// idxT / depthT / distsT [K][B]: entry (k, r) is defined for k < lens[r] only.  A CTA takes kTR = 128 rays and walks
// their slots in tiles of 32: each warp reads 32 consecutive slots of 8 rays (128-byte row pieces), the tile is turned
// in shared memory, and each warp writes 2 slots for the 128 rays (512 contiguous bytes per slot and plane, so the
// write stream keeps some DRAM page locality although consecutive slots of a plane lie 4*B bytes apart).
constexpr int kTR = 128;
__global__ void __launch_bounds__(512)
march_transpose_kernel(long long B, long long ldb, int K, long long ldk, int k_begin, int k_end,
                       const unsigned char* __restrict__ early_stop, const int* __restrict__ lens, const int* __restrict__ idx,
                       const float* __restrict__ depth, const float* __restrict__ dists, int* __restrict__ idxT,
                       float* __restrict__ depthT, float* __restrict__ distsT) {
  extern __shared__ int tr_smem[];
  int (*t_i)[33] = reinterpret_cast<int (*)[33]>(tr_smem);
  float (*t_d)[33] = reinterpret_cast<float (*)[33]>(tr_smem + kTR * 33);
  float (*t_s)[33] = reinterpret_cast<float (*)[33]>(tr_smem + 2 * kTR * 33);
  int* s_len = tr_smem + 3 * kTR * 33;
  __shared__ int s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 16 warps
  for (long long r0 = (long long)blockIdx.x * kTR; r0 < B; r0 += (long long)gridDim.x * kTR) {
    __syncthreads();
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (threadIdx.x < kTR) {
      int l = (r0 + threadIdx.x < B) ? lens[r0 + threadIdx.x] : 0;
      l = l < 0 ? 0 : (l > k_end ? k_end : l);
      if (l > 0 && early_stop != nullptr && early_stop[r0 + threadIdx.x]) l = 0;   // stopped rays need no more columns
      s_len[threadIdx.x] = l;
      int m = l;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(NSVF_FULL_MASK, m, o));
      if (lane == 0) atomicMax(&s_max, m);
    }
    __syncthreads();
    const int maxlen = s_max;
    for (int k0 = k_begin; k0 < maxlen; k0 += 32) {
#pragma unroll
      for (int j = 0; j < kTR / 16; ++j) {
        const int rr = warp * (kTR / 16) + j;
        const int k = k0 + lane;
        if (k < s_len[rr]) {
          const long long a = (r0 + rr) * ldk + k;
          t_i[rr][lane] = idx[a];
          t_d[rr][lane] = depth[a];
          t_s[rr][lane] = dists[a];
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kk = warp * 2 + j;
        const int k = k0 + kk;
#pragma unroll
        for (int q = 0; q < kTR / 32; ++q) {
          const int rr = q * 32 + lane;
          if (k < s_len[rr]) {
            const long long a = (long long)k * ldb + r0 + rr;
            idxT[a] = t_i[rr][kk];
            depthT[a] = t_d[rr][kk];
            distsT[a] = t_s[rr][kk];
          }
        }
      }
      __syncthreads();
    }
  }
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

// The window of a launch queued ahead of its read-back (start < 0) lives in the plan header: ONE thread per CTA reads it.
__device__ __forceinline__ void resolve_window(const int* __restrict__ plan, int& start, int& end) {
  __shared__ int sh_window[2];
  if (start >= 0) return;          // uniform across the CTA (kernel argument)
  if (threadIdx.x == 0) {
    const int s0 = __ldcg(plan + H_START);
    sh_window[0] = s0;
    sh_window[1] = __ldcg(plan + H_DONE) ? s0 : __ldcg(plan + H_END);
  }
  __syncthreads();
  start = sh_window[0];
  end = sh_window[1];
}

__device__ __forceinline__ int block_sum(int x) {      // sum over the CTA (kTile threads), result in every thread
  __shared__ int sh_part[kTile / 32], sh_sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = __reduce_add_sync(NSVF_FULL_MASK, x);
  if (lane == 0) sh_part[warp] = w;
  __syncthreads();
  if (warp == 0) {
    const int t = __reduce_add_sync(NSVF_FULL_MASK, lane < kTile / 32 ? sh_part[lane] : 0);
    if (lane == 0) sh_sum = t;
  }
  __syncthreads();
  return sh_sum;
}

// pass 1: samples of every 256-ray tile in the window
__global__ void __launch_bounds__(kTile)
march_count_kernel(long long B, int start, int end, const int* __restrict__ lens,
                   const unsigned char* __restrict__ early_stop, const int* __restrict__ plan,
                   int* __restrict__ tile_total) {
  resolve_window(plan, start, end);
  const long long ray = (long long)blockIdx.x * kTile + threadIdx.x;
  int n = 0;
  if (ray < B && (early_stop == nullptr || early_stop[ray] == 0)) n = window_samples(lens[ray], start, end);
  const int total = block_sum(n);
  if (threadIdx.x == 0) tile_total[blockIdx.x] = total;
}

// pass 2: offsets and samples
__global__ void __launch_bounds__(kTile)
march_compact_kernel(long long B, long long ldb, int K, int start, int end, const int* __restrict__ lens,
                     const unsigned char* __restrict__ early_stop, const int* __restrict__ idxT,
                     const float* __restrict__ depthT, const float* __restrict__ distsT,
                     const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                     int* __restrict__ out_vox, float* __restrict__ out_xyz, float* __restrict__ out_dir,
                     float* __restrict__ out_dists, int* __restrict__ ray_off, const int* __restrict__ plan,
                     const int* __restrict__ tile_total) {
  __shared__ int sh_warp[kTile / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  resolve_window(plan, start, end);
  const long long ray = (long long)blockIdx.x * kTile + tid;
  int n = 0;
  if (ray < B && (early_stop == nullptr || early_stop[ray] == 0)) n = window_samples(lens[ray], start, end);
  // the first column's loads do not depend on the offsets: issue them before the sums
  int v0 = 0;
  float d0 = 0.f, s0 = 0.f;
  if (n > 0) {
    const long long a = (long long)start * ldb + ray;
    v0 = idxT[a]; d0 = depthT[a]; s0 = distsT[a];
  }
  // samples of all tiles before this one
  int part = 0;
  for (int i = tid; i < (int)blockIdx.x; i += kTile) part += __ldcg(tile_total + i);
  const int base = block_sum(part);
  // exclusive scan of n inside the tile
  const int incl = warp_incl_sum_i(n, lane);
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int w = 0; w < kTile / 32; ++w) wbase += w < warp ? sh_warp[w] : 0;
  const int off = base + wbase + incl - n;
  if (ray < B) ray_off[ray] = off;
  if (ray == B - 1) ray_off[B] = off + n;
  if (n > 0) {
    const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
    const float dx = ray_dir[ray * 3 + 0], dy = ray_dir[ray * 3 + 1], dz = ray_dir[ray * 3 + 2];
    for (int t = 0; t < n; ++t) {
      if (t > 0) {
        const long long a = (long long)(start + t) * ldb + ray;
        v0 = idxT[a]; d0 = depthT[a]; s0 = distsT[a];
      }
      const long long o = (long long)off + t;
      out_vox[o] = v0;
      // ray(): ray_start + ray_dir * depth — a separate multiply and add in the reference (no FMA)
      out_xyz[o * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, d0));
      out_xyz[o * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, d0));
      out_xyz[o * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, d0));
      if (out_dir != nullptr) { out_dir[o * 3 + 0] = dx; out_dir[o * 3 + 1] = dy; out_dir[o * 3 + 2] = dz; }
      if (out_dists != nullptr) out_dists[o] = s0;
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
march_epilogue_kernel(long long B, long long ldb, int K, int start, int end, const int* __restrict__ ray_off,
                      const int* __restrict__ lens, unsigned char* __restrict__ early_stop,
                      float* __restrict__ acc_fe, int* __restrict__ eval_len, const float* __restrict__ sigma,
                      const float* __restrict__ noise, const float* __restrict__ dists,
                      const float* __restrict__ texture, float tolerance, float* __restrict__ feT,
                      float* __restrict__ texT, int chunk_size, int schedule_next, int* __restrict__ plan,
                      volatile int* host_info) {
  const int tid = threadIdx.x;
  int* diff = plan + kHdr + K;
  // every ray that stops in this window takes one live sample off the columns [end, len): the -1's all land on
  // diff[end] — thousands of atomics on ONE address per window when done per ray — so they are summed per CTA first
  __shared__ int sh_stopped;
  if (tid == 0) sh_stopped = 0;
  __syncthreads();
  for (long long ray = (long long)blockIdx.x * kTile + tid; ray < B; ray += (long long)gridDim.x * kTile) {
    const int o0 = ray_off[ray], n = ray_off[ray + 1] - o0;
    if (n == 0) continue;
    float sum = 0.f;
    for (int t = 0; t < n; ++t) {
      const long long o = (long long)o0 + t;
      const long long a = (long long)(start + t) * ldb + ray;
      if (sigma != nullptr) {
        const float fe = free_energy(sigma[o], noise != nullptr ? noise[o] : 0.f, dists[o]);
        feT[a] = fe;
        sum += fe;
      }
      if (texture != nullptr) {
        texT[a * 3 + 0] = texture[o * 3 + 0];
        texT[a * 3 + 1] = texture[o * 3 + 1];
        texT[a * 3 + 2] = texture[o * 3 + 2];
      }
    }
    eval_len[ray] = start + n;
    if (tolerance > 0.f && sigma != nullptr) {
      const float acc = acc_fe[ray] + sum;
      acc_fe[ray] = acc;
      if (acc > tolerance) {
        early_stop[ray] = 1;
        const int len = lens[ray];
        if (len > end) { atomicAdd(&sh_stopped, 1); atomicAdd(diff + len, 1); }
      }
    }
  }
  __syncthreads();
  if (tid == 0 && sh_stopped != 0) atomicSub(diff + end, sh_stopped);
  if (schedule_next && last_cta(plan) && tid < 32)
    publish_schedule(plan, K, chunk_size, end, true, false, host_info, 0);
}

// backward of the epilogue for one window: gradients of the planes back to the compacted field outputs
//   d sigma = (g_fe * 7) * dists * [noise + sigma > 0];  d texture = g_tex
__global__ void __launch_bounds__(256)
march_epilogue_bwd_kernel(long long B, long long ldb, int K, int start, const int* __restrict__ ray_off,
                          const float* __restrict__ g_feT, const float* __restrict__ g_texT,
                          const float* __restrict__ sigma, const float* __restrict__ noise,
                          const float* __restrict__ dists, float* __restrict__ g_sigma, float* __restrict__ g_texture) {
  for (long long ray = (long long)blockIdx.x * 256 + threadIdx.x; ray < B; ray += (long long)gridDim.x * 256) {
    const int o0 = ray_off[ray], n = ray_off[ray + 1] - o0;
    for (int t = 0; t < n; ++t) {
      const long long o = (long long)o0 + t;
      const long long a = (long long)(start + t) * ldb + ray;
      if (g_sigma != nullptr) {
        const float x = __fadd_rn(noise != nullptr ? noise[o] : 0.f, sigma[o]);
        g_sigma[o] = x > 0.f ? __fmul_rn(__fmul_rn(g_feT[a], 7.0f), dists[o]) : 0.f;
      }
      if (g_texture != nullptr) {
        g_texture[o * 3 + 0] = g_texT[a * 3 + 0];
        g_texture[o * 3 + 1] = g_texT[a * 3 + 1];
        g_texture[o * 3 + 2] = g_texT[a * 3 + 2];
      }
    }
  }
}

// ---- compositing over the slot-major planes (renderer.py:193-218), one thread per ray ------------------------------
//   a = 1 - exp(-fe);  b = exp(-cumsum(shift(fe)));  probs = a * b;  depth / missed / colors = sums over the ray
// The running free-energy sum is compensated (Kahan): it is the one long sequential sum of the scan.
__global__ void __launch_bounds__(256)
march_composite_fwd_kernel(long long B, long long ldb, int K, const int* __restrict__ eval_len, const int* __restrict__ lens,
                           const unsigned char* __restrict__ early_stop, const float* __restrict__ feT,
                           const float* __restrict__ texT, const float* __restrict__ depthT,
                           float* __restrict__ probsT, float* __restrict__ out_depth, float* __restrict__ out_missed,
                           float* __restrict__ out_colors, float* __restrict__ out_maxd, float* __restrict__ out_mind,
                           const float* __restrict__ depth_rows, long long ldk, float pad_depth, int lazy_planes) {
  const long long ray = (long long)blockIdx.x * 256 + threadIdx.x;
  if (ray >= B) return;
  const int Lfull = min(max(lens[ray], 0), K);
  const int Kr = min(eval_len[ray], Lfull);
  // lazily transposed planes hold only the evaluated prefix of a ray that stopped early; its depth extrema do not
  // need more: max_depths is -1 for it and the minimum of a depth-ordered ray is its first sample
  const int Lr = (lazy_planes && early_stop != nullptr && early_stop[ray]) ? Kr : Lfull;
  float carry = 0.f, comp = 0.f, s_p = 0.f, s_d = 0.f, s_r = 0.f, s_g = 0.f, s_b = 0.f;
  float dmax = -1.0f, dmin = 3.0e38f;
  int k = 0;
  for (; k + 4 <= Kr; k += 4) {                     // 4 columns in flight per thread
    float x[4], d[4], t[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long a = (long long)(k + j) * ldb + ray;
      x[j] = feT[a];
      d[j] = depthT[a];
      if (texT != nullptr) { t[j][0] = texT[a * 3 + 0]; t[j][1] = texT[a * 3 + 1]; t[j][2] = texT[a * 3 + 2]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = (1.0f - expf(-x[j])) * expf(-carry);
      const float y = x[j] - comp, tt = carry + y;
      comp = (tt - carry) - y;
      carry = tt;
      if (probsT != nullptr) probsT[(long long)(k + j) * ldb + ray] = p;
      s_p += p;
      s_d = fmaf(d[j], p, s_d);
      if (texT != nullptr) { s_r = fmaf(t[j][0], p, s_r); s_g = fmaf(t[j][1], p, s_g); s_b = fmaf(t[j][2], p, s_b); }
      dmax = fmaxf(dmax, d[j]);
      dmin = fminf(dmin, d[j]);
    }
  }
  for (; k < Lr; ++k) {
    const long long a = (long long)k * ldb + ray;
    const float dk = depthT[a];
    dmax = fmaxf(dmax, dk);
    dmin = fminf(dmin, dk);
    float p = 0.f;
    if (k < Kr) {
      const float xk = feT[a];
      p = (1.0f - expf(-xk)) * expf(-carry);
      const float y = xk - comp, tt = carry + y;
      comp = (tt - carry) - y;
      carry = tt;
      s_p += p;
      s_d = fmaf(dk, p, s_d);
      if (texT != nullptr) {
        s_r = fmaf(texT[a * 3 + 0], p, s_r); s_g = fmaf(texT[a * 3 + 1], p, s_g); s_b = fmaf(texT[a * 3 + 2], p, s_b);
      }
    }
    if (probsT != nullptr) probsT[a] = p;
  }
  if (probsT != nullptr)
    for (; k < K; ++k) probsT[(long long)k * ldb + ray] = 0.f;
  out_depth[ray] = s_d;
  out_missed[ray] = 1.0f - s_p;
  if (texT != nullptr && out_colors != nullptr) {
    out_colors[ray * 3 + 0] = s_r; out_colors[ray * 3 + 1] = s_g; out_colors[ray * 3 + 2] = s_b;
  }
  // renderer.py:210-211: max over the live samples (-1 for rays that stopped early / have none), min over the whole
  // padded row (padding = pad_depth, or the first padding slot of a padded input row)
  if (out_maxd != nullptr) out_maxd[ray] = (early_stop != nullptr && early_stop[ray]) ? -1.0f : dmax;
  if (out_mind != nullptr) {
    if (Lfull < K) dmin = fminf(dmin, depth_rows != nullptr ? depth_rows[ray * ldk + Lfull] : pad_depth);
    out_mind[ray] = dmin;
  }
}

// backward, two sweeps per ray:  G_k = dprobs_k + ddepth * t_k - dmissed + dcolors . rgb_k
//   d fe_k = G_k e^{-fe_k} b_k - sum_{j>k} G_j probs_j ;  d rgb_k = dcolors * probs_k
// sweep 1 (ascending) stores the first term in g_feT and q_k = G_k probs_k in scratchT; sweep 2 (descending)
// subtracts the exclusive suffix sum — a true reverse scan, no cancellation for late samples.
__global__ void __launch_bounds__(256)
march_composite_bwd_kernel(long long B, long long ldb, int K, const int* __restrict__ eval_len, const float* __restrict__ feT,
                           const float* __restrict__ texT, const float* __restrict__ depthT,
                           const float* __restrict__ g_probsT, const float* __restrict__ g_depth,
                           const float* __restrict__ g_missed, const float* __restrict__ g_colors,
                           float* __restrict__ g_feT, float* __restrict__ g_texT, float* __restrict__ scratchT) {
  const long long ray = (long long)blockIdx.x * 256 + threadIdx.x;
  if (ray >= B) return;
  const int Kr = min(eval_len[ray], K);
  if (Kr == 0) return;
  const float gd = g_depth ? g_depth[ray] : 0.f, gm = g_missed ? g_missed[ray] : 0.f;
  float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
  if (g_colors && texT) { gc0 = g_colors[ray * 3 + 0]; gc1 = g_colors[ray * 3 + 1]; gc2 = g_colors[ray * 3 + 2]; }
  float carry = 0.f, comp = 0.f;
  for (int k = 0; k < Kr; ++k) {
    const long long a = (long long)k * ldb + ray;
    const float x = feT[a];
    const float e = expf(-x), bk = expf(-carry);
    const float y = x - comp, tt = carry + y;
    comp = (tt - carry) - y;
    carry = tt;
    const float p = (1.0f - e) * bk;
    float G = (g_probsT ? g_probsT[a] : 0.f) + gd * depthT[a] - gm;
    if (texT) {
      G += gc0 * texT[a * 3 + 0] + gc1 * texT[a * 3 + 1] + gc2 * texT[a * 3 + 2];
      if (g_texT) { g_texT[a * 3 + 0] = gc0 * p; g_texT[a * 3 + 1] = gc1 * p; g_texT[a * 3 + 2] = gc2 * p; }
    }
    g_feT[a] = G * e * bk;
    scratchT[a] = G * p;
  }
  float tail = 0.f;
  for (int k = Kr - 1; k >= 0; --k) {
    const long long a = (long long)k * ldb + ray;
    g_feT[a] -= tail;
    tail += scratchT[a];
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" long long nsvf_march_plane_stride(long long B) { return B < 0 ? 0 : plane_stride(B); }

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

extern "C" int nsvf_march_transpose(nsvf_stream_t stream_, long long B, int K, long long ldk, int k_begin, int k_end,
                                    const unsigned char* early_stop, const int* lens, const int* sampled_idx, const float* sampled_depth, const float* sampled_dists,
                                    int* idxT, float* depthT, float* distsT) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && ldk >= K && k_begin >= 0 && k_begin <= k_end && k_end <= K && k_begin % 32 == 0,
               "march_transpose: bad sizes (k_begin must be a multiple of 32)");
  if (B == 0 || k_end == k_begin) return 0;
  const size_t smem = (size_t)(3 * kTR * 33 + kTR) * sizeof(int);
  NSVF_CUDA_OK(cudaFuncSetAttribute(march_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long want = (B + kTR - 1) / kTR, cap = (long long)num_sms() * 4;
  NSVF_TIMED_LAUNCH("march_transpose_kernel", stream,
                    (march_transpose_kernel<<<(int)(want < cap ? want : cap), 512, smem, stream>>>(
                        B, plane_stride(B), K, ldk, k_begin, k_end, early_stop, lens, sampled_idx, sampled_depth, sampled_dists, idxT, depthT, distsT)));
  return 0;
}

extern "C" int nsvf_march_compact(nsvf_stream_t stream_, long long B, int K, int start, int end, const int* lens,
                                  const unsigned char* early_stop, const int* idxT, const float* depthT,
                                  const float* distsT, const float* ray_start, const float* ray_dir, int* out_vox,
                                  float* out_xyz, float* out_dir, float* out_dists, int* ray_off, void* plan,
                                  int launch_no) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && (start == -1 || (start >= 0 && start <= end && end <= K)) && launch_no >= 0,
               "march_compact: bad sizes");
  if (B == 0) return 0;
  (void)launch_no;   // (the first-generation single-pass scan tagged its tile states with the launch number)
  const long long tiles = plan_tiles(B);
  int* tile_total = (int*)plan + plan_words_before_tiles(K);
  profile_mark("march_compact_kernel", 0, stream);      // the two passes are timed as one unit
  march_count_kernel<<<(unsigned)tiles, kTile, 0, stream>>>(B, start, end, lens, early_stop, (const int*)plan, tile_total);
  NSVF_LAUNCH_OK("march_count_kernel");
  march_compact_kernel<<<(unsigned)tiles, kTile, 0, stream>>>(
      B, plane_stride(B), K, start, end, lens, early_stop, idxT, depthT, distsT, ray_start, ray_dir, out_vox, out_xyz, out_dir,
      out_dists, ray_off, (const int*)plan, tile_total);
  profile_mark("march_compact_kernel", 1, stream);
  NSVF_LAUNCH_OK("march_compact_kernel");
  return 0;
}

extern "C" int nsvf_march_epilogue(nsvf_stream_t stream_, long long B, int K, int start, int end, const int* ray_off,
                                   const int* lens, unsigned char* early_stop, float* acc_free_energy, int* eval_len,
                                   const float* sigma, const float* noise, const float* dists, const float* texture,
                                   float tolerance, float* feT, float* texT, int chunk_size, int schedule_next,
                                   void* plan, int* host_info) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && start >= 0 && start <= end && end <= K, "march_epilogue: bad sizes");
  NSVF_REQUIRE(sigma == nullptr || (dists != nullptr && feT != nullptr), "march_epilogue: sigma needs dists");
  NSVF_REQUIRE(texture == nullptr || texT != nullptr, "march_epilogue: texture needs its plane");
  if (B == 0) return 0;
  long long want = (B + kTile - 1) / kTile, cap = (long long)num_sms() * 8;
  NSVF_TIMED_LAUNCH("march_epilogue_kernel", stream,
                    (march_epilogue_kernel<<<(int)(want < cap ? want : cap), kTile, 0, stream>>>(
                        B, plane_stride(B), K, start, end, ray_off, lens, early_stop, acc_free_energy, eval_len, sigma, noise, dists,
                        texture, tolerance, feT, texT, chunk_size, schedule_next, (int*)plan, host_info)));
  return 0;
}

extern "C" int nsvf_march_epilogue_bwd(nsvf_stream_t stream_, long long B, int K, int start, int end,
                                       const int* ray_off, const float* g_feT, const float* g_texT, const float* sigma,
                                       const float* noise, const float* dists, float* grad_sigma,
                                       float* grad_texture) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && start >= 0 && start <= end && end <= K, "march_epilogue_bwd: bad sizes");
  if (B == 0) return 0;
  long long want = (B + 255) / 256, cap = (long long)num_sms() * 8;
  march_epilogue_bwd_kernel<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(
      B, plane_stride(B), K, start, ray_off, g_feT, g_texT, sigma, noise, dists, grad_sigma, grad_texture);
  NSVF_LAUNCH_OK("march_epilogue_bwd_kernel");
  return 0;
}

extern "C" int nsvf_march_composite_fwd(nsvf_stream_t stream_, long long B, int K, const int* eval_len,
                                        const int* lens, const unsigned char* early_stop, const float* feT,
                                        const float* texT, const float* depthT, float* probsT, float* depth,
                                        float* missed, float* colors, float* max_depths, float* min_depths,
                                        const float* padded_depth_rows, long long ldk, float pad_depth,
                                        int lazy_planes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0, "march_composite_fwd: bad sizes");
  NSVF_REQUIRE(eval_len != nullptr && lens != nullptr, "march_composite_fwd: eval_len and lens are required");
  if (B == 0) return 0;
  NSVF_TIMED_LAUNCH("march_composite_fwd_kernel", stream,
                    (march_composite_fwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(
                        B, plane_stride(B), K, eval_len, lens, early_stop, feT, texT, depthT, probsT, depth, missed, colors, max_depths,
                        min_depths, padded_depth_rows, ldk, pad_depth, lazy_planes)));
  return 0;
}

extern "C" int nsvf_march_composite_bwd(nsvf_stream_t stream_, long long B, int K, const int* eval_len,
                                        const float* feT, const float* texT, const float* depthT,
                                        const float* grad_probsT, const float* grad_depth, const float* grad_missed,
                                        const float* grad_colors, float* g_feT, float* g_texT, float* scratchT) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0, "march_composite_bwd: bad sizes");
  NSVF_REQUIRE(eval_len != nullptr && scratchT != nullptr, "march_composite_bwd: eval_len and scratch are required");
  if (B == 0 || K == 0) return 0;
  NSVF_TIMED_LAUNCH("march_composite_bwd_kernel", stream,
                    (march_composite_bwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(
                        B, plane_stride(B), K, eval_len, feT, texT, depthT, grad_probsT, grad_depth, grad_missed, grad_colors, g_feT,
                        g_texT, scratchT)));
  return 0;
}
