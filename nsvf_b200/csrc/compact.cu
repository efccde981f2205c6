// Device-side sample compaction for the renderer: replaces the boolean-mask indexing of
// VolumeRenderer.forward_once, fairnr/modules/renderer.py:88-100 (sample_mask = idx != -1 & ~early_stop;
// xyz = ray_start + ray_dir * depth; {name: s[sample_mask]}) — five torch.nonzero-style compactions,
// each with a host sync — by a count kernel + a fill kernel that emit the compacted voxel id, position,
// direction, step length and the flat [B,K] position of every valid sample in row-major order
// (the order boolean indexing produces), with no host synchronisation.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

constexpr int kCompactWarps = 8;

__global__ void __launch_bounds__(kCompactWarps * 32)
compact_count_kernel(long long B, int K, int col0, int col1, const int* __restrict__ sampled_idx,
                     const unsigned char* __restrict__ early_stop, long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  for (long long ray = (long long)blockIdx.x * kCompactWarps + (threadIdx.x >> 5); ray < B;
       ray += (long long)gridDim.x * kCompactWarps) {
    int c = 0;
    if (early_stop == nullptr || early_stop[ray] == 0) {
      for (int k = col0 + lane; k < col1; k += 32) c += (sampled_idx[ray * K + k] != -1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(NSVF_FULL_MASK, c, o);
    if (lane == 0) counts[ray] = c;
  }
}

__global__ void __launch_bounds__(kCompactWarps * 32)
compact_fill_kernel(long long B, int K, int col0, int col1, const int* __restrict__ sampled_idx,
                    const float* __restrict__ sampled_depth, const float* __restrict__ sampled_dists,
                    const unsigned char* __restrict__ early_stop, const float* __restrict__ ray_start,
                    const float* __restrict__ ray_dir, const long long* __restrict__ offsets_incl,
                    int* __restrict__ out_vox, float* __restrict__ out_xyz, float* __restrict__ out_dir,
                    float* __restrict__ out_dists, long long* __restrict__ out_flat) {
  const int lane = threadIdx.x & 31;
  for (long long ray = (long long)blockIdx.x * kCompactWarps + (threadIdx.x >> 5); ray < B;
       ray += (long long)gridDim.x * kCompactWarps) {
    if (early_stop != nullptr && early_stop[ray] != 0) continue;
    long long base = ray == 0 ? 0 : offsets_incl[ray - 1];   // inclusive prefix sums of the counts
    const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
    const float dx = ray_dir[ray * 3 + 0], dy = ray_dir[ray * 3 + 1], dz = ray_dir[ray * 3 + 2];
    for (int k0 = col0; k0 < col1; k0 += 32) {
      const int k = k0 + lane;
      int v = -1;
      if (k < col1) v = sampled_idx[ray * K + k];
      const bool ok = v != -1;
      const unsigned m = __ballot_sync(NSVF_FULL_MASK, ok);
      if (ok) {
        const long long o = base + __popc(m & ((1u << lane) - 1u));
        const float t = sampled_depth[ray * K + k];
        out_vox[o] = v;
        // ray(): ray_start + ray_dir * depth, a separate multiply and add in the reference (no FMA)
        out_xyz[o * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, t));
        out_xyz[o * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, t));
        out_xyz[o * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, t));
        if (out_dir != nullptr) { out_dir[o * 3 + 0] = dx; out_dir[o * 3 + 1] = dy; out_dir[o * 3 + 2] = dz; }
        if (out_dists != nullptr) out_dists[o] = sampled_dists[ray * K + k];
        out_flat[o] = ray * K + k;
      }
      base += __popc(m);
    }
  }
}

// Narrow column windows (the usual case: a chunk spans a handful of sample columns): one THREAD per ray.
__global__ void __launch_bounds__(256)
compact_count_narrow_kernel(long long B, int K, int col0, int col1, const int* __restrict__ sampled_idx,
                            const unsigned char* __restrict__ early_stop, long long* __restrict__ counts) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B) return;
  int c = 0;
  if (early_stop == nullptr || early_stop[ray] == 0)
    for (int k = col0; k < col1; ++k) c += (sampled_idx[ray * K + k] != -1);
  counts[ray] = c;
}

__global__ void __launch_bounds__(256)
compact_fill_narrow_kernel(long long B, int K, int col0, int col1, const int* __restrict__ sampled_idx,
                           const float* __restrict__ sampled_depth, const float* __restrict__ sampled_dists,
                           const unsigned char* __restrict__ early_stop, const float* __restrict__ ray_start,
                           const float* __restrict__ ray_dir, const long long* __restrict__ offsets_incl,
                           int* __restrict__ out_vox, float* __restrict__ out_xyz, float* __restrict__ out_dir,
                           float* __restrict__ out_dists, long long* __restrict__ out_flat) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B) return;
  if (early_stop != nullptr && early_stop[ray] != 0) return;
  const long long end = offsets_incl[ray];
  long long o = ray == 0 ? 0 : offsets_incl[ray - 1];
  if (o == end) return;
  const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
  const float dx = ray_dir[ray * 3 + 0], dy = ray_dir[ray * 3 + 1], dz = ray_dir[ray * 3 + 2];
  for (int k = col0; k < col1; ++k) {
    const int v = sampled_idx[ray * K + k];
    if (v == -1) continue;
    const float t = sampled_depth[ray * K + k];
    out_vox[o] = v;
    out_xyz[o * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, t));
    out_xyz[o * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, t));
    out_xyz[o * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, t));
    if (out_dir != nullptr) { out_dir[o * 3 + 0] = dx; out_dir[o * 3 + 1] = dy; out_dir[o * 3 + 2] = dz; }
    if (out_dists != nullptr) out_dists[o] = sampled_dists[ray * K + k];
    out_flat[o] = ray * K + k;
    ++o;
  }
}

constexpr int kCompactNarrow = 12;   // windows up to this many columns use the thread-per-ray kernels

// Per-column counts of valid samples among rays that have not stopped: counts[k - col0] = #{ray : idx[ray,k] != -1
// and !early_stop[ray]} for k in [col0, col1) — the quantity the reference's chunk scheduler reads with one
// `hits[:, i].sum()` host sync per column (renderer.py:158,187).  A CTA takes 256 rays; each warp sweeps its 32
// rays x the window with lanes along the columns (coalesced) and flushes one shared-memory atomic per column,
// the CTA one global atomic per column.
__global__ void __launch_bounds__(256)
masked_col_counts_kernel(long long B, int K, int col0, int col1, const int* __restrict__ sampled_idx,
                         const unsigned char* __restrict__ early_stop, int* __restrict__ counts) {
  extern __shared__ int col_smem[];
  const int ncol = col1 - col0;
  for (int k = threadIdx.x; k < ncol; k += blockDim.x) col_smem[k] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * 256 + warp * 32;
  const long long r1 = min(B, r0 + 32);
  for (int k0 = 0; k0 < ncol; k0 += 32) {
    const int k = k0 + lane;
    int c = 0;
    if (k < ncol) {
      for (long long ray = r0; ray < r1; ++ray) {
        if (early_stop != nullptr && early_stop[ray]) continue;
        c += (sampled_idx[ray * K + col0 + k] != -1);
      }
      if (c) atomicAdd(col_smem + k, c);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ncol; k += blockDim.x)
    if (col_smem[k]) atomicAdd(counts + k, col_smem[k]);
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_masked_col_counts(nsvf_stream_t stream_, long long B, int K, int col0, int col1,
                                      const int* sampled_idx, const unsigned char* early_stop, int* counts) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && col0 >= 0 && col0 <= col1 && col1 <= K, "masked_col_counts: bad sizes");
  if (col1 == col0) return 0;
  NSVF_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)(col1 - col0), stream));
  if (B == 0) return 0;
  const size_t smem = sizeof(int) * (size_t)(col1 - col0);
  NSVF_REQUIRE(smem <= 48 * 1024, "masked_col_counts: window too wide");
  masked_col_counts_kernel<<<(unsigned)((B + 255) / 256), 256, smem, stream>>>(B, K, col0, col1, sampled_idx,
                                                                               early_stop, counts);
  NSVF_LAUNCH_OK("masked_col_counts_kernel");
  return 0;
}

extern "C" int nsvf_compact_count(nsvf_stream_t stream_, long long B, int K, int col0, int col1,
                                  const int* sampled_idx, const unsigned char* early_stop, long long* counts) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && col0 >= 0 && col1 >= col0 && col1 <= K, "compact_count: bad sizes");
  if (B == 0) return 0;
  if (col1 - col0 <= kCompactNarrow) {
    compact_count_narrow_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(B, K, col0, col1, sampled_idx,
                                                                               early_stop, counts);
    NSVF_LAUNCH_OK("compact_count_narrow_kernel");
    return 0;
  }
  long long want = (B + kCompactWarps - 1) / kCompactWarps, cap = (long long)num_sms() * 8;
  compact_count_kernel<<<(int)(want < cap ? want : cap), kCompactWarps * 32, 0, stream>>>(B, K, col0, col1, sampled_idx,
                                                                                      early_stop, counts);
  NSVF_LAUNCH_OK("compact_count_kernel");
  return 0;
}

extern "C" int nsvf_compact_fill(nsvf_stream_t stream_, long long B, int K, int col0, int col1,
                                 const int* sampled_idx, const float* sampled_depth, const float* sampled_dists,
                                 const unsigned char* early_stop, const float* ray_start, const float* ray_dir,
                                 const long long* offsets_incl, int* out_vox, float* out_xyz, float* out_dir,
                                 float* out_dists, long long* out_flat) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && col0 >= 0 && col1 >= col0 && col1 <= K, "compact_fill: bad sizes");
  if (B == 0) return 0;
  if (col1 - col0 <= kCompactNarrow) {
    compact_fill_narrow_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(
        B, K, col0, col1, sampled_idx, sampled_depth, sampled_dists, early_stop, ray_start, ray_dir, offsets_incl,
        out_vox, out_xyz, out_dir, out_dists, out_flat);
    NSVF_LAUNCH_OK("compact_fill_narrow_kernel");
    return 0;
  }
  long long want = (B + kCompactWarps - 1) / kCompactWarps, cap = (long long)num_sms() * 8;
  compact_fill_kernel<<<(int)(want < cap ? want : cap), kCompactWarps * 32, 0, stream>>>(
      B, K, col0, col1, sampled_idx, sampled_depth, sampled_dists, early_stop, ray_start, ray_dir, offsets_incl,
      out_vox, out_xyz, out_dir, out_dists, out_flat);
  NSVF_LAUNCH_OK("compact_fill_kernel");
  return 0;
}
