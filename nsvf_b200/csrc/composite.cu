// Alpha compositing along rays (forward + fused backward) for sm_100a.
//
// Replaces VolumeRenderer.forward_chunk's compositing block, fairnr/modules/renderer.py:193-218:
//   a = 1 - exp(-fe);  b = exp(-cumsum(shift(fe)));  probs = a * b
//   depth = sum(t * probs);  missed = 1 - sum(probs);  colors = sum(rgb * probs)
// which the reference runs as >= 8 elementwise / scan / reduce kernels over [B,K] (plus autograd's
// saved intermediates).  Here one warp owns one ray: lanes stride over the K samples (coalesced), the
// exclusive prefix of free energy is a warp shuffle scan with a running carry, and the three
// reductions share the same pass.  The backward recomputes probs from fe (nothing saved but the
// inputs) and uses a true reverse (suffix) scan, so late samples do not suffer cancellation:
//   G_k      = dL/dprobs_k + dL/ddepth * t_k - dL/dmissed + dL/dcolors . rgb_k
//   dL/dfe_k = G_k * exp(-fe_k) * b_k  -  sum_{j>k} G_j * probs_j
//   dL/drgb_k = dL/dcolors * probs_k
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

constexpr int kCompWarps = 8;

__device__ __forceinline__ float warp_incl_scan(float x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(NSVF_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(NSVF_FULL_MASK, x, o);
  return x;
}

// Trimmed rows (march.cu): eval_len[ray] = number of leading samples whose free energy / texture exist (the rest
// count as zero free energy), lens[ray] = number of leading samples whose depth exists; both NULL = dense rows.
// fe / tex / probs rows have stride K, depth rows stride ldk.
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(long long B, int K, long long ldk, const int* __restrict__ eval_len,
                     const int* __restrict__ lens, const unsigned char* __restrict__ early_stop,
                     const float* __restrict__ fe, const float* __restrict__ tex,
                     const float* __restrict__ depth, float* __restrict__ probs, float* __restrict__ out_depth,
                     float* __restrict__ out_missed, float* __restrict__ out_colors, float* __restrict__ out_maxd,
                     float* __restrict__ out_mind, float pad_depth, int padded) {
  const int lane = threadIdx.x & 31;
  for (long long ray = (long long)blockIdx.x * kCompWarps + (threadIdx.x >> 5); ray < B;
       ray += (long long)gridDim.x * kCompWarps) {
    const long long row = ray * K, drow = ray * ldk;
    const int Lr = lens != nullptr ? min(lens[ray], K) : K;
    const int Kr = eval_len != nullptr ? min(eval_len[ray], Lr) : Lr;
    float carry = 0.f, s_p = 0.f, s_d = 0.f, s_r = 0.f, s_g = 0.f, s_b = 0.f;
    float dmax = -1.0f, dmin = 3.0e38f;
    // software pipeline: chunk k0+32's inputs are in flight while chunk k0 is scanned
    float nx = 0.f, nd = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (lane < Lr) nd = depth[drow + lane];
    if (lane < Kr) {
      nx = fe[row + lane];
      if (tex != nullptr) { n0 = tex[(row + lane) * 3 + 0]; n1 = tex[(row + lane) * 3 + 1]; n2 = tex[(row + lane) * 3 + 2]; }
    }
    for (int k0 = 0; k0 < Lr; k0 += 32) {
      const int k = k0 + lane;
      const bool ok = k < Kr;
      const float x = ok ? nx : 0.f, dk = nd, t0 = n0, t1 = n1, t2 = n2;
      const int kn = k + 32;
      if (kn < Lr) nd = depth[drow + kn];
      if (kn < Kr) {
        nx = fe[row + kn];
        if (tex != nullptr) { n0 = tex[(row + kn) * 3 + 0]; n1 = tex[(row + kn) * 3 + 1]; n2 = tex[(row + kn) * 3 + 2]; }
      }
      if (k < Lr) { dmax = fmaxf(dmax, dk); dmin = fminf(dmin, dk); }
      if (k0 >= Kr) {                      // beyond the evaluated prefix: zero free energy, zero probability
        if (probs != nullptr && k < K) probs[row + k] = 0.f;
        continue;
      }
      const float incl = warp_incl_scan(x, lane);
      float excl = __shfl_up_sync(NSVF_FULL_MASK, incl, 1);
      if (lane == 0) excl = 0.f;
      excl += carry;
      carry += __shfl_sync(NSVF_FULL_MASK, incl, 31);
      if (ok) {
        const float p = (1.0f - expf(-x)) * expf(-excl);
        if (probs != nullptr) probs[row + k] = p;
        s_p += p;
        s_d = fmaf(dk, p, s_d);
        if (tex != nullptr) {
          s_r = fmaf(t0, p, s_r);
          s_g = fmaf(t1, p, s_g);
          s_b = fmaf(t2, p, s_b);
        }
      } else if (probs != nullptr && k < K) {
        probs[row + k] = 0.f;
      }
    }
    if (probs != nullptr)                 // rest of the dense probs row
      for (int k = ((Lr + 31) / 32) * 32 + lane; k < K; k += 32) probs[row + k] = 0.f;
    s_p = warp_sum(s_p);
    s_d = warp_sum(s_d);
    if (tex != nullptr) { s_r = warp_sum(s_r); s_g = warp_sum(s_g); s_b = warp_sum(s_b); }
    if (out_maxd != nullptr || out_mind != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dmax = fmaxf(dmax, __shfl_xor_sync(NSVF_FULL_MASK, dmax, o));
        dmin = fminf(dmin, __shfl_xor_sync(NSVF_FULL_MASK, dmin, o));
      }
    }
    if (lane == 0) {
      out_depth[ray] = s_d;
      out_missed[ray] = 1.0f - s_p;
      if (tex != nullptr && out_colors != nullptr) {
        out_colors[ray * 3 + 0] = s_r;
        out_colors[ray * 3 + 1] = s_g;
        out_colors[ray * 3 + 2] = s_b;
      }
      // renderer.py:210-211: max over the live samples (-1 for rays that stopped early / have none), min over the
      // whole padded row (padding = pad_depth, or whatever the first padding slot of a padded tensor holds)
      if (out_maxd != nullptr) out_maxd[ray] = (early_stop != nullptr && early_stop[ray]) ? -1.0f : dmax;
      if (out_mind != nullptr) {
        if (Lr < K) dmin = fminf(dmin, padded ? depth[drow + Lr] : pad_depth);
        out_mind[ray] = dmin;
      }
    }
  }
}

__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(long long B, int K, long long ldk, const int* __restrict__ eval_len,
                     const float* __restrict__ fe, const float* __restrict__ tex,
                     const float* __restrict__ depth, const float* __restrict__ g_probs,
                     const float* __restrict__ g_depth, const float* __restrict__ g_missed,
                     const float* __restrict__ g_colors, float* __restrict__ g_fe, float* __restrict__ g_tex) {
  extern __shared__ float comp_smem[];  // per warp: first[K], q[K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* first = comp_smem + (size_t)warp * 2 * K;
  float* q = first + K;
  for (long long ray = (long long)blockIdx.x * kCompWarps + warp; ray < B; ray += (long long)gridDim.x * kCompWarps) {
    const long long row = ray * K, drow = ray * ldk;
    const int Kr = eval_len != nullptr ? min(eval_len[ray], K) : K;   // gradients exist for the evaluated prefix only
    if (Kr == 0) continue;
    const float gd = g_depth ? g_depth[ray] : 0.f;
    const float gm = g_missed ? g_missed[ray] : 0.f;
    float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
    if (g_colors && tex) { gc0 = g_colors[ray * 3 + 0]; gc1 = g_colors[ray * 3 + 1]; gc2 = g_colors[ray * 3 + 2]; }
    // pass 1 (forward): probs, G, first term, q = G * probs.  Inputs of chunk k0+32 are loaded while chunk k0 is
    // scanned (the carry makes the chunks sequential, the loads need not be).
    float carry = 0.f;
    float nx = 0.f, nd = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, ngp = 0.f;
    if (lane < Kr) {
      nx = fe[row + lane];
      nd = depth[drow + lane];
      if (tex) { n0 = tex[(row + lane) * 3 + 0]; n1 = tex[(row + lane) * 3 + 1]; n2 = tex[(row + lane) * 3 + 2]; }
      if (g_probs) ngp = g_probs[row + lane];
    }
    for (int k0 = 0; k0 < Kr; k0 += 32) {
      const int k = k0 + lane;
      const bool ok = k < Kr;
      const float x = ok ? nx : 0.f, dk = nd, t0 = n0, t1 = n1, t2 = n2, gp = ngp;
      const int kn = k + 32;
      if (kn < Kr) {
        nx = fe[row + kn];
        nd = depth[drow + kn];
        if (tex) { n0 = tex[(row + kn) * 3 + 0]; n1 = tex[(row + kn) * 3 + 1]; n2 = tex[(row + kn) * 3 + 2]; }
        if (g_probs) ngp = g_probs[row + kn];
      }
      const float incl = warp_incl_scan(x, lane);
      float excl = __shfl_up_sync(NSVF_FULL_MASK, incl, 1);
      if (lane == 0) excl = 0.f;
      excl += carry;
      carry += __shfl_sync(NSVF_FULL_MASK, incl, 31);
      if (ok) {
        const float e = expf(-x), bk = expf(-excl);
        const float p = (1.0f - e) * bk;
        float G = gp + gd * dk - gm;
        if (tex) {
          G += gc0 * t0 + gc1 * t1 + gc2 * t2;
          if (g_tex) {
            g_tex[(row + k) * 3 + 0] = gc0 * p;
            g_tex[(row + k) * 3 + 1] = gc1 * p;
            g_tex[(row + k) * 3 + 2] = gc2 * p;
          }
        }
        first[k] = G * e * bk;
        q[k] = G * p;
      }
    }
    __syncwarp();
    // pass 2 (reverse): exclusive suffix sum of q
    float tail = 0.f;
    for (int k0 = ((Kr - 1) / 32) * 32; k0 >= 0; k0 -= 32) {
      const int k = k0 + (31 - lane);  // lane 0 takes the last element of the chunk
      const bool ok = k < Kr;
      const float x = ok ? q[k] : 0.f;
      const float incl = warp_incl_scan(x, lane);
      float excl = __shfl_up_sync(NSVF_FULL_MASK, incl, 1);
      if (lane == 0) excl = 0.f;
      excl += tail;
      tail += __shfl_sync(NSVF_FULL_MASK, incl, 31);
      if (ok) g_fe[row + k] = first[k] - excl;
    }
    __syncwarp();
  }
}

}  // namespace nsvf

using namespace nsvf;

static int composite_fwd_launch(cudaStream_t stream, long long B, int K, long long ldk, const int* eval_len,
                                const int* lens, const unsigned char* early_stop, const float* free_energy,
                                const float* texture, const float* sampled_depth, float* probs, float* depth,
                                float* missed, float* colors, float* max_depths, float* min_depths, float pad_depth,
                                int padded) {
  long long want = (B + kCompWarps - 1) / kCompWarps;
  long long cap = (long long)num_sms() * 8;
  int grid = (int)(want < cap ? want : cap);
  NSVF_TIMED_LAUNCH("composite_fwd_kernel", stream,
                    (composite_fwd_kernel<<<grid, kCompWarps * 32, 0, stream>>>(
                        B, K, ldk, eval_len, lens, early_stop, free_energy, texture, sampled_depth, probs, depth,
                        missed, colors, max_depths, min_depths, pad_depth, padded)));
  return 0;
}

static int composite_bwd_launch(cudaStream_t stream, long long B, int K, long long ldk, const int* eval_len,
                                const float* free_energy, const float* texture, const float* sampled_depth,
                                const float* grad_probs, const float* grad_depth, const float* grad_missed,
                                const float* grad_colors, float* grad_free_energy, float* grad_texture) {
  const size_t smem = (size_t)kCompWarps * 2 * K * sizeof(float);
  NSVF_REQUIRE(smem <= 200 * 1024, "composite_bwd: K=%d too large for the shared-memory stash (%zu B)", K, smem);
  NSVF_CUDA_OK(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long want = (B + kCompWarps - 1) / kCompWarps;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  long long cap = (long long)num_sms() * per_sm;
  int grid = (int)(want < cap ? want : cap);
  NSVF_TIMED_LAUNCH("composite_bwd_kernel", stream,
                    (composite_bwd_kernel<<<grid, kCompWarps * 32, smem, stream>>>(
                        B, K, ldk, eval_len, free_energy, texture, sampled_depth, grad_probs, grad_depth, grad_missed,
                        grad_colors, grad_free_energy, grad_texture)));
  return 0;
}

extern "C" int nsvf_composite_fwd(nsvf_stream_t stream_, long long B, int K, const float* free_energy,
                                  const float* texture, const float* sampled_depth, float* probs, float* depth,
                                  float* missed, float* colors) {
  NSVF_REQUIRE(B >= 0 && K >= 0, "composite_fwd: negative size");
  if (B == 0) return 0;
  return composite_fwd_launch((cudaStream_t)stream_, B, K, K, nullptr, nullptr, nullptr, free_energy, texture,
                              sampled_depth, probs, depth, missed, colors, nullptr, nullptr, 0.f, 1);
}

extern "C" int nsvf_composite_bwd(nsvf_stream_t stream_, long long B, int K, const float* free_energy,
                                  const float* texture, const float* sampled_depth, const float* grad_probs,
                                  const float* grad_depth, const float* grad_missed, const float* grad_colors,
                                  float* grad_free_energy, float* grad_texture) {
  NSVF_REQUIRE(B >= 0 && K >= 0, "composite_bwd: negative size");
  if (B == 0 || K == 0) return 0;
  return composite_bwd_launch((cudaStream_t)stream_, B, K, K, nullptr, free_energy, texture, sampled_depth, grad_probs,
                              grad_depth, grad_missed, grad_colors, grad_free_energy, grad_texture);
}

