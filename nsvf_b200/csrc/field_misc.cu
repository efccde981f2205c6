// The non-GEMM ends of the field MLP, sm_100a:
//   * NeRF positional encoding (fairnr/modules/module_utils.py:56-87, NeRFPosEmbLinear with no_linear): torch runs it as
//     an outer product, sin, cos, two cats and a clone — eight passes over [M, 384] floats; here one pass writes the
//     [M, C*2L (+C)] row, and the backward is one pass over the incoming gradient.
//   * the two output heads (nn.Linear 128 -> 1 for sigma, 256 -> 3 for rgb): with 1 or 3 output features these are
//     not contractions a tensor core can use — cuBLAS runs them as SIMT GEMM / GEMV launches (130 us for the weight
//     gradient alone); here they are streaming kernels bound by reading x once.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

// ---- positional encoding ----------------------------------------------------------------------------------------
// out[s, c*2L + k] = sin(f_k * t), out[s, c*2L + L + k] = cos(f_k * t), t = x[s,c] or acos(clamp(x[s,c])) (angular),
// optionally followed by the C raw inputs.  One thread per (sample, channel).
template <int L>
__global__ void __launch_bounds__(256)
posenc_fwd_kernel(long long M, int C, const float* __restrict__ x, const float* __restrict__ freq, int angular,
                  int cat_input, float* __restrict__ out) {
  const int row = C * 2 * L + (cat_input ? C : 0);
  float f[L];
#pragma unroll
  for (int k = 0; k < L; ++k) f[k] = __ldg(freq + k);
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / C;
    const int c = (int)(i - s * C);
    const float xin = __ldg(x + i);
    const float t = angular ? acosf(fminf(fmaxf(xin, -1.0f + 1e-6f), 1.0f - 1e-6f)) : xin;
    float* o = out + s * row + c * 2 * L;
    float v[2 * L];   // [sin(f_0 t) .. sin(f_{L-1} t), cos(f_0 t) .. cos(f_{L-1} t)]
#pragma unroll
    for (int k = 0; k < L; ++k) sincosf(__fmul_rn(t, f[k]), &v[k], &v[L + k]);
    if ((2 * L) % 4 == 0 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int k = 0; k < 2 * L; k += 4) reinterpret_cast<float4*>(o)[k >> 2] = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 2 * L; ++k) o[k] = v[k];
    }
    if (cat_input) out[s * row + C * 2 * L + c] = xin;
  }
}

// d x[s,c] = sum_k f_k * (cos(f_k x) * g_sin[k] - sin(f_k x) * g_cos[k]) (+ g_x[c] when cat_input); non-angular only
template <int L>
__global__ void __launch_bounds__(256)
posenc_bwd_kernel(long long M, int C, const float* __restrict__ x, const float* __restrict__ freq, int cat_input,
                  const float* __restrict__ grad_out, float* __restrict__ grad_x) {
  const int row = C * 2 * L + (cat_input ? C : 0);
  float f[L];
#pragma unroll
  for (int k = 0; k < L; ++k) f[k] = __ldg(freq + k);
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / C;
    const int c = (int)(i - s * C);
    const float t = __ldg(x + i);
    const float* g = grad_out + s * row + c * 2 * L;
    float gv[2 * L];
    if ((2 * L) % 4 == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
#pragma unroll
      for (int k = 0; k < 2 * L; k += 4) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(g) + (k >> 2));
        gv[k] = a.x; gv[k + 1] = a.y; gv[k + 2] = a.z; gv[k + 3] = a.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 2 * L; ++k) gv[k] = __ldcs(g + k);
    }
    const float* gs = gv;
    const float* gc = gv + L;
    float acc = cat_input ? __ldcs(grad_out + s * row + C * 2 * L + c) : 0.f;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      float sn, cs;
      sincosf(__fmul_rn(t, f[k]), &sn, &cs);
      acc = fmaf(f[k], fmaf(cs, gs[k], -sn * gc[k]), acc);
    }
    grad_x[i] = acc;
  }
}

// ---- narrow linear: y[M, O] = x[M, K] W[O, K]^T + b, O <= 4, K in {128, 256, 512} -----------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSVF_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

template <int K, int O>
__global__ void __launch_bounds__(256)
narrow_linear_fwd_kernel(long long M, const float* __restrict__ x, const float* __restrict__ W,
                         const float* __restrict__ b, float* __restrict__ y) {
  constexpr int V = K / 128;
  const int lane = threadIdx.x & 31;
  float4 w[O][V];
#pragma unroll
  for (int o = 0; o < O; ++o)
#pragma unroll
    for (int k = 0; k < V; ++k) w[o][k] = __ldg(reinterpret_cast<const float4*>(W + (long long)o * K) + lane + 32 * k);
  float bias[O];
#pragma unroll
  for (int o = 0; o < O; ++o) bias[o] = b != nullptr ? __ldg(b + o) : 0.f;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < M; r += nwarps) {
    float4 xv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) xv[k] = __ldg(reinterpret_cast<const float4*>(x + r * K) + lane + 32 * k);
    float acc[O];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) a += dot4(xv[k], w[o][k]);
      acc[o] = warp_sum_f(a) + bias[o];
    }
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < O; ++o) y[r * O + o] = acc[o];
    }
  }
}

// dx[M, K] = dy[M, O] W[O, K]; partial[grid][O*K + O]: per-CTA sums of dy^T x (dW) and of dy (db)
template <int K, int O>
__global__ void __launch_bounds__(256)
narrow_linear_bwd_kernel(long long M, const float* __restrict__ x, const float* __restrict__ W,
                         const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ partial) {
  constexpr int V = K / 128;
  constexpr int RS = (O * K + O + 3) & ~3;   // row stride in floats, 16-byte aligned rows
  __shared__ __align__(16) float red[8][RS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 w[O][V], aw[O][V];
  float ab[O];
#pragma unroll
  for (int o = 0; o < O; ++o) {
    ab[o] = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      w[o][k] = __ldg(reinterpret_cast<const float4*>(W + (long long)o * K) + lane + 32 * k);
      aw[o][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const long long warp0 = (long long)blockIdx.x * 8 + warp, nwarps = (long long)gridDim.x * 8;
  for (long long r = warp0; r < M; r += nwarps) {
    float4 xv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) xv[k] = __ldcs(reinterpret_cast<const float4*>(x + r * K) + lane + 32 * k);
    float g[O];
#pragma unroll
    for (int o = 0; o < O; ++o) g[o] = __ldg(dy + r * O + o);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < O; ++o) {
        d.x = fmaf(g[o], w[o][k].x, d.x); d.y = fmaf(g[o], w[o][k].y, d.y);
        d.z = fmaf(g[o], w[o][k].z, d.z); d.w = fmaf(g[o], w[o][k].w, d.w);
        aw[o][k].x = fmaf(g[o], xv[k].x, aw[o][k].x); aw[o][k].y = fmaf(g[o], xv[k].y, aw[o][k].y);
        aw[o][k].z = fmaf(g[o], xv[k].z, aw[o][k].z); aw[o][k].w = fmaf(g[o], xv[k].w, aw[o][k].w);
      }
      if (dx != nullptr) reinterpret_cast<float4*>(dx + r * K)[lane + 32 * k] = d;
    }
#pragma unroll
    for (int o = 0; o < O; ++o) ab[o] += g[o];
  }
#pragma unroll
  for (int o = 0; o < O; ++o) {
#pragma unroll
    for (int k = 0; k < V; ++k) reinterpret_cast<float4*>(red[warp] + o * K)[lane + 32 * k] = aw[o][k];
    if (lane == 0) red[warp][O * K + o] = ab[o];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O * K + O; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) s += red[wi][i];
    partial[(long long)blockIdx.x * (O * K + O) + i] = s;
  }
}

// out[c] = sum_b partial[b][c], fixed order; columns [0, n0) go to o0, the rest to o1
__global__ void __launch_bounds__(1024)
colsum_partials_kernel(int blocks, int cols, const float* __restrict__ partial, int n0, float* __restrict__ o0,
                       float* __restrict__ o1) {
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
    if (col < n0) {
      if (o0 != nullptr) o0[col] = t;
    } else if (o1 != nullptr) {
      o1[col - n0] = t;
    }
  }
}

static int rows_grid(long long M, int cap_per_sm) {
  long long want = (M + 7) / 8, cap = (long long)num_sms() * cap_per_sm;
  long long g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_posenc_fwd(nsvf_stream_t stream_, long long M, int C, int L, const float* x, const float* freq,
                               int angular, int cat_input, float* out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0 && C > 0 && (L == 4 || L == 6 || L == 10), "posenc_fwd: L must be 4, 6 or 10");
  if (M == 0) return 0;
  const long long total = M * C;
  const int grid = (int)((total + 255) / 256 < (long long)num_sms() * 16 ? (total + 255) / 256 : (long long)num_sms() * 16);
  if (L == 4) {
    NSVF_TIMED_LAUNCH("posenc_fwd_kernel", stream, (posenc_fwd_kernel<4><<<grid, 256, 0, stream>>>(M, C, x, freq, angular, cat_input, out)));
  } else if (L == 6) {
    NSVF_TIMED_LAUNCH("posenc_fwd_kernel", stream, (posenc_fwd_kernel<6><<<grid, 256, 0, stream>>>(M, C, x, freq, angular, cat_input, out)));
  } else {
    NSVF_TIMED_LAUNCH("posenc_fwd_kernel", stream, (posenc_fwd_kernel<10><<<grid, 256, 0, stream>>>(M, C, x, freq, angular, cat_input, out)));
  }
  return 0;
}

extern "C" int nsvf_posenc_bwd(nsvf_stream_t stream_, long long M, int C, int L, const float* x, const float* freq,
                               int cat_input, const float* grad_out, float* grad_x) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0 && C > 0 && (L == 4 || L == 6 || L == 10), "posenc_bwd: L must be 4, 6 or 10");
  if (M == 0) return 0;
  const long long total = M * C;
  const int grid = (int)((total + 255) / 256 < (long long)num_sms() * 16 ? (total + 255) / 256 : (long long)num_sms() * 16);
  if (L == 4) {
    NSVF_TIMED_LAUNCH("posenc_bwd_kernel", stream, (posenc_bwd_kernel<4><<<grid, 256, 0, stream>>>(M, C, x, freq, cat_input, grad_out, grad_x)));
  } else if (L == 6) {
    NSVF_TIMED_LAUNCH("posenc_bwd_kernel", stream, (posenc_bwd_kernel<6><<<grid, 256, 0, stream>>>(M, C, x, freq, cat_input, grad_out, grad_x)));
  } else {
    NSVF_TIMED_LAUNCH("posenc_bwd_kernel", stream, (posenc_bwd_kernel<10><<<grid, 256, 0, stream>>>(M, C, x, freq, cat_input, grad_out, grad_x)));
  }
  return 0;
}

#define NSVF_NARROW_DISPATCH(MACRO)                                     \
  if (K == 128 && O == 1) { MACRO(128, 1); }                            \
  else if (K == 128 && O == 3) { MACRO(128, 3); }                       \
  else if (K == 256 && O == 1) { MACRO(256, 1); }                       \
  else if (K == 256 && O == 3) { MACRO(256, 3); }                       \
  else if (K == 256 && O == 4) { MACRO(256, 4); }                       \
  else { nsvf::set_error("narrow_linear: unsupported (K, O) = (%d, %d)", K, O); return 1; }

extern "C" int nsvf_narrow_linear_supported(int K, int O) {
  return ((K == 128 && (O == 1 || O == 3)) || (K == 256 && (O == 1 || O == 3 || O == 4))) ? 1 : 0;
}

extern "C" int nsvf_narrow_linear_fwd(nsvf_stream_t stream_, long long M, int K, int O, const float* x, const float* W,
                                      const float* b, float* y) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0, "narrow_linear_fwd: bad M");
  NSVF_REQUIRE((((uintptr_t)x | (uintptr_t)W) & 15) == 0, "narrow_linear_fwd: x / W must be 16-byte aligned");
  if (M == 0) return 0;
  const int grid = rows_grid(M, 8);
#define NSVF_NL_FWD(KK, OO)                                                                                      \
  NSVF_TIMED_LAUNCH("narrow_linear_fwd_kernel", stream,                                                          \
                    (narrow_linear_fwd_kernel<KK, OO><<<grid, 256, 0, stream>>>(M, x, W, b, y)))
  NSVF_NARROW_DISPATCH(NSVF_NL_FWD)
#undef NSVF_NL_FWD
  return 0;
}

extern "C" size_t nsvf_narrow_linear_bwd_workspace_bytes(long long M, int K, int O) {
  return (size_t)rows_grid(M, 4) * (size_t)(O * K + O) * sizeof(float);
}

extern "C" int nsvf_narrow_linear_bwd(nsvf_stream_t stream_, long long M, int K, int O, const float* x, const float* W,
                                      const float* dy, float* dx, float* dW, float* db, void* workspace,
                                      size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M > 0, "narrow_linear_bwd: M > 0 required");
  NSVF_REQUIRE((((uintptr_t)x | (uintptr_t)W | (uintptr_t)dx | (uintptr_t)workspace) & 15) == 0,
               "narrow_linear_bwd: pointers must be 16-byte aligned");
  const int grid = rows_grid(M, 4);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)grid * (O * K + O) * sizeof(float),
               "narrow_linear_bwd: workspace too small (nsvf_narrow_linear_bwd_workspace_bytes)");
  float* partial = static_cast<float*>(workspace);
#define NSVF_NL_BWD(KK, OO)                                                                                      \
  NSVF_TIMED_LAUNCH("narrow_linear_bwd_kernel", stream,                                                          \
                    (narrow_linear_bwd_kernel<KK, OO><<<grid, 256, 0, stream>>>(M, x, W, dy, dx, partial)))
  NSVF_NARROW_DISPATCH(NSVF_NL_BWD)
#undef NSVF_NL_BWD
  const int cols = O * K + O;
  NSVF_TIMED_LAUNCH("colsum_partials_kernel", stream,
                    (colsum_partials_kernel<<<(cols + 31) / 32, 1024, 0, stream>>>(grid, cols, partial, O * K, dW, db)));
  return 0;
}
