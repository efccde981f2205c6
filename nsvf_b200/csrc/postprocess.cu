// Per-frame post-processing on the hot path.
//
// nsvf_fill_in_blend  — replaces fill_in (fairnr/data/geometry.py:303-317) x3 + the background blend of
//   NSVFModel.postprocessing (fairnr/models/nsvf.py:89-104): results of the hit rays are scattered into full-size
//   images and `missed * bg_color` / `missed * BG_DEPTH` are added, in one pass (reference: 3 masked_scatter +
//   ~6 elementwise kernels over the full image).
// nsvf_track_voxel_probs — replaces SparseVoxelEncoder.track_voxel_probs (fairnr/modules/encoder.py:594-603), which
//   allocates a [4096, n+1] scatter_add buffer per 4096-ray chunk (1.6 GB at n = 100 k): per ray, the probabilities a
//   ray deposits in one voxel are consecutive samples, so one pass with a running sum and an atomicMax suffices.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

__global__ void fill_in_blend_kernel(long long N, const unsigned char* __restrict__ hits,
                                     const long long* __restrict__ rank_incl, const float* __restrict__ colors,
                                     const float* __restrict__ missed, const float* __restrict__ depths,
                                     const float* __restrict__ bg, float bg_depth, float* __restrict__ out_colors,
                                     float* __restrict__ out_missed, float* __restrict__ out_depths) {
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, m = 1.0f, d = 0.f;   // fill_in defaults: colors 0, missed 1, depths 0
    if (hits[i]) {
      const long long j = rank_incl[i] - 1;
      c0 = colors[j * 3 + 0]; c1 = colors[j * 3 + 1]; c2 = colors[j * 3 + 2];
      m = missed[j];
      d = depths[j];
    }
    // all_results['colors'] += missed * bg_color ; all_results['depths'] += missed * BG_DEPTH  (mul, then add)
    out_colors[i * 3 + 0] = __fadd_rn(c0, __fmul_rn(m, bg0));
    out_colors[i * 3 + 1] = __fadd_rn(c1, __fmul_rn(m, bg1));
    out_colors[i * 3 + 2] = __fadd_rn(c2, __fmul_rn(m, bg2));
    out_missed[i] = m;
    out_depths[i] = __fadd_rn(d, __fmul_rn(m, bg_depth));
  }
}

__global__ void track_voxel_probs_kernel(long long B, int K, const int* __restrict__ sampled_idx,
                                         const float* __restrict__ probs, int n_vox, float* __restrict__ max_probs) {
  for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < B;
       ray += (long long)gridDim.x * blockDim.x) {
    int cur = -1;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const int v = sampled_idx[ray * K + k];
      if (v != cur) {
        if (cur >= 0 && cur < n_vox)   // probs >= 0: the int order of the bit patterns is the float order
          atomicMax(reinterpret_cast<int*>(max_probs + cur), __float_as_int(fmaxf(acc, 0.0f)));
        cur = v;
        acc = 0.f;
      }
      if (v >= 0) acc += probs[ray * K + k];
    }
    if (cur >= 0 && cur < n_vox)
      atomicMax(reinterpret_cast<int*>(max_probs + cur), __float_as_int(fmaxf(acc, 0.0f)));
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_fill_in_blend(nsvf_stream_t stream_, long long N, const unsigned char* hits,
                                  const long long* rank_incl, const float* colors, const float* missed,
                                  const float* depths, const float* bg_color, float bg_depth, float* out_colors,
                                  float* out_missed, float* out_depths) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(N >= 0 && bg_color != nullptr, "fill_in_blend: bad arguments");
  if (N == 0) return 0;
  long long want = (N + 255) / 256, cap = (long long)num_sms() * 16;
  fill_in_blend_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(
      N, hits, rank_incl, colors, missed, depths, bg_color, bg_depth, out_colors, out_missed, out_depths);
  NSVF_LAUNCH_OK("fill_in_blend_kernel");
  return 0;
}

extern "C" int nsvf_track_voxel_probs(nsvf_stream_t stream_, long long B, int K, const int* sampled_idx,
                                      const float* probs, int n_vox, float* max_probs) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(B >= 0 && K >= 0 && n_vox >= 0, "track_voxel_probs: negative size");
  if (B == 0 || K == 0 || n_vox == 0) return 0;
  long long want = (B + 127) / 128, cap = (long long)num_sms() * 16;
  track_voxel_probs_kernel<<<(unsigned)(want < cap ? want : cap), 128, 0, stream>>>(B, K, sampled_idx, probs, n_vox,
                                                                                  max_probs);
  NSVF_LAUNCH_OK("track_voxel_probs_kernel");
  return 0;
}
