// Fused LayerNorm + ReLU of the field MLP's FCLayer (fairnr/modules/module_utils.py:97-111: Linear -> LayerNorm([o])
// -> ReLU) and its backward, sm_100a.  The Linear itself is a dense contraction and stays on cuBLAS; what is fused
// here is everything around it that torch runs as separate [M, N] passes:
//   forward :  layer_norm + affine + relu                         (3 passes  -> 1: read h, write y)
//   backward:  relu mask, d gamma, d beta, layer-norm input grad and the Linear's bias gradient (column sum of dh)
//              (torch: threshold_backward, 3 column reductions of [M, N] at 80 us each, layer_norm_grad_input, 3
//              elementwise products -> 1 pass: read h and dy, write dh, plus one tiny partial-sum reduction)
// One warp per row; a lane holds N/32 values of the row in registers (float4 at columns 4*lane + 128*k), so every
// global access is a full 512-byte warp transaction.  Column sums are accumulated per lane in registers across the
// rows a warp walks, reduced per CTA through shared memory, written as per-CTA partials and summed by a second
// kernel in a fixed order (deterministic, no atomics).
//
//   xhat = (h - mean) * rstd,  pre = xhat * gamma + beta,  y = max(pre, 0)
//   g    = dy * [pre > 0];  d beta = sum_rows g;  d gamma = sum_rows g * xhat
//   dh   = rstd * (g*gamma - mean_cols(g*gamma) - xhat * mean_cols(g*gamma*xhat));  d bias = sum_rows dh
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSVF_FULL_MASK, v, o);
  return v;
}

// the one expression both directions use for the pre-activation, so the ReLU mask of the backward is the forward's
__device__ __forceinline__ float ln_pre(float x, float mean, float rstd, float gamma, float beta, float& xhat) {
  xhat = __fmul_rn(__fsub_rn(x, mean), rstd);
  return __fmaf_rn(xhat, gamma, beta);
}

template <int N>
__global__ void __launch_bounds__(256)
ln_relu_fwd_kernel(long long M, const float* __restrict__ h, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                   float* __restrict__ rstd_out) {
  constexpr int V = N / 128;
  const int lane = threadIdx.x & 31;
  float4 gm[V], bt[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    gm[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    bt[k] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * k);
  }
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < M; r += nwarps) {
    const float4* row = reinterpret_cast<const float4*>(h + r * N);
    float4 x[V];
#pragma unroll
    for (int k = 0; k < V; ++k) x[k] = __ldg(row + lane + 32 * k);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
    const float mean = warp_sum(s) * (1.0f / N);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float a = x[k].x - mean, b = x[k].y - mean, c = x[k].z - mean, d = x[k].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / N) + eps);
    float4* out = reinterpret_cast<float4*>(y + r * N);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float xh;
      float4 o;
      o.x = fmaxf(ln_pre(x[k].x, mean, rstd, gm[k].x, bt[k].x, xh), 0.f);
      o.y = fmaxf(ln_pre(x[k].y, mean, rstd, gm[k].y, bt[k].y, xh), 0.f);
      o.z = fmaxf(ln_pre(x[k].z, mean, rstd, gm[k].z, bt[k].z, xh), 0.f);
      o.w = fmaxf(ln_pre(x[k].w, mean, rstd, gm[k].w, bt[k].w, xh), 0.f);
      out[lane + 32 * k] = o;   // default policy: the next Linear reads y straight away (L2)
    }
    if (lane == 0) {
      if (mean_out != nullptr) mean_out[r] = mean;
      if (rstd_out != nullptr) rstd_out[r] = rstd;
    }
  }
}

// partial: f32 [gridDim.x][3][N]  (d gamma, d beta, d bias)
template <int N>
__global__ void __launch_bounds__(256)
ln_relu_bwd_kernel(long long M, const float* __restrict__ h, const float* __restrict__ dy,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, float* __restrict__ dh, float* __restrict__ partial) {
  constexpr int V = N / 128;
  __shared__ float red[8][3][N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 gm[V], bt[V], a_g[V], a_b[V], a_h[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    gm[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    bt[k] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * k);
    a_g[k] = a_b[k] = a_h[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long warp0 = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long r = warp0; r < M; r += nwarps) {
    const float4* hrow = reinterpret_cast<const float4*>(h + r * N);
    const float4* grow = reinterpret_cast<const float4*>(dy + r * N);
    float4 x[V], g[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      x[k] = __ldcs(hrow + lane + 32 * k);   // last use of h and dy: streaming
      g[k] = __ldcs(grow + lane + 32 * k);
    }
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      // x <- xhat, g <- g * gamma (masked); accumulate d gamma / d beta on the way
#define NSVF_LN_ELEM(c)                                                        \
      {                                                                        \
        float xh;                                                              \
        const float pre = ln_pre(x[k].c, mean, rstd, gm[k].c, bt[k].c, xh);    \
        const float gz = pre > 0.f ? g[k].c : 0.f;                             \
        a_b[k].c += gz;                                                        \
        a_g[k].c = fmaf(gz, xh, a_g[k].c);                                     \
        const float gx = gz * gm[k].c;                                         \
        s1 += gx;                                                              \
        s2 = fmaf(gx, xh, s2);                                                 \
        x[k].c = xh;                                                           \
        g[k].c = gx;                                                           \
      }
      NSVF_LN_ELEM(x) NSVF_LN_ELEM(y) NSVF_LN_ELEM(z) NSVF_LN_ELEM(w)
#undef NSVF_LN_ELEM
    }
    s1 = warp_sum(s1) * (1.0f / N);
    s2 = warp_sum(s2) * (1.0f / N);
    float4* out = reinterpret_cast<float4*>(dh + r * N);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 o;
      o.x = rstd * (g[k].x - s1 - x[k].x * s2);
      o.y = rstd * (g[k].y - s1 - x[k].y * s2);
      o.z = rstd * (g[k].z - s1 - x[k].z * s2);
      o.w = rstd * (g[k].w - s1 - x[k].w * s2);
      a_h[k].x += o.x; a_h[k].y += o.y; a_h[k].z += o.z; a_h[k].w += o.w;
      out[lane + 32 * k] = o;   // read next by the two cuBLAS GEMMs (dW, dx)
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    reinterpret_cast<float4*>(red[warp][0])[lane + 32 * k] = a_g[k];
    reinterpret_cast<float4*>(red[warp][1])[lane + 32 * k] = a_b[k];
    reinterpret_cast<float4*>(red[warp][2])[lane + 32 * k] = a_h[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * N; i += 256) {
    const int which = i / N, c = i - which * N;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][which][c];
    partial[(long long)blockIdx.x * 3 * N + i] = s;
  }
}

// out[c] = sum_b partial[b][c] for the 3N columns; one CTA per 32 columns, 32 row groups x 32 columns, fixed
// summation order (deterministic).
__global__ void __launch_bounds__(1024)
ln_partial_sum_kernel(int blocks, int cols, const float* __restrict__ partial, float* __restrict__ o0,
                      float* __restrict__ o1, float* __restrict__ o2, int N) {
  __shared__ float red[32][33];
  const int c = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + c;
  float s = 0.f;
  if (col < cols) {
#pragma unroll 4
    for (int b = j; b < blocks; b += 32) s += partial[(long long)b * cols + col];
  }
  red[j][c] = s;
  __syncthreads();
  if (j == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][c];
    const int which = col / N, cc = col - which * N;
    float* o = which == 0 ? o0 : (which == 1 ? o1 : o2);
    if (o != nullptr) o[cc] = t;
  }
}

static int ln_grid(long long M, int cap_per_sm) {
  long long want = (M + 7) / 8;
  long long cap = (long long)num_sms() * cap_per_sm;
  long long g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_ln_relu_fwd(nsvf_stream_t stream_, long long M, int N, const float* h, const float* gamma,
                                const float* beta, float eps, float* y, float* mean, float* rstd) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0 && (N == 128 || N == 256 || N == 512), "ln_relu_fwd: N must be 128, 256 or 512");
  NSVF_REQUIRE((((uintptr_t)h | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
               "ln_relu_fwd: pointers must be 16-byte aligned");
  if (M == 0) return 0;
  const int grid = ln_grid(M, 8);
  if (N == 128) {
    NSVF_TIMED_LAUNCH("ln_relu_fwd_kernel", stream,
                      (ln_relu_fwd_kernel<128><<<grid, 256, 0, stream>>>(M, h, gamma, beta, eps, y, mean, rstd)));
  } else if (N == 256) {
    NSVF_TIMED_LAUNCH("ln_relu_fwd_kernel", stream,
                      (ln_relu_fwd_kernel<256><<<grid, 256, 0, stream>>>(M, h, gamma, beta, eps, y, mean, rstd)));
  } else {
    NSVF_TIMED_LAUNCH("ln_relu_fwd_kernel", stream,
                      (ln_relu_fwd_kernel<512><<<grid, 256, 0, stream>>>(M, h, gamma, beta, eps, y, mean, rstd)));
  }
  return 0;
}

extern "C" size_t nsvf_ln_relu_bwd_workspace_bytes(long long M, int N) {
  return (size_t)ln_grid(M, 4) * 3 * (size_t)N * sizeof(float);
}

extern "C" int nsvf_ln_relu_bwd(nsvf_stream_t stream_, long long M, int N, const float* h, const float* dy,
                                const float* gamma, const float* beta, const float* mean, const float* rstd,
                                float* dh, float* dgamma, float* dbeta, float* dbias, void* workspace,
                                size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M > 0 && (N == 128 || N == 256 || N == 512), "ln_relu_bwd: M > 0 and N in {128, 256, 512} required");
  NSVF_REQUIRE((((uintptr_t)h | (uintptr_t)dy | (uintptr_t)dh | (uintptr_t)gamma | (uintptr_t)beta |
                 (uintptr_t)workspace) & 15) == 0, "ln_relu_bwd: pointers must be 16-byte aligned");
  const int grid = ln_grid(M, 4);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)grid * 3 * N * sizeof(float),
               "ln_relu_bwd: workspace too small (nsvf_ln_relu_bwd_workspace_bytes)");
  float* partial = static_cast<float*>(workspace);
  if (N == 128) {
    NSVF_TIMED_LAUNCH("ln_relu_bwd_kernel", stream,
                      (ln_relu_bwd_kernel<128><<<grid, 256, 0, stream>>>(M, h, dy, gamma, beta, mean, rstd, dh, partial)));
  } else if (N == 256) {
    NSVF_TIMED_LAUNCH("ln_relu_bwd_kernel", stream,
                      (ln_relu_bwd_kernel<256><<<grid, 256, 0, stream>>>(M, h, dy, gamma, beta, mean, rstd, dh, partial)));
  } else {
    NSVF_TIMED_LAUNCH("ln_relu_bwd_kernel", stream,
                      (ln_relu_bwd_kernel<512><<<grid, 256, 0, stream>>>(M, h, dy, gamma, beta, mean, rstd, dh, partial)));
  }
  NSVF_TIMED_LAUNCH("ln_partial_sum_kernel", stream,
                    (ln_partial_sum_kernel<<<(3 * N + 31) / 32, 1024, 0, stream>>>(grid, 3 * N, partial, dgamma, dbeta,
                                                                                   dbias, N)));
  return 0;
}
