// C ABI of the Linear contractions of the field MLP (the three GEMMs of an FCLayer) on the CUTLASS BF16x9 kernels of
// csrc/gemm/ (see gemm/fast_gemm.cuh for what they are and why they replace the cuBLAS call).
#include <cstdlib>

#include "common.cuh"
#include "nsvf_b200.h"

#ifndef NSVF_NO_CUTLASS
#define NSVF_GEMM_DECL(name)                                                                                          \
  extern "C" int name(int M, int N, int K, int L, const float* A, long long a_batch, const float* B, long long b_batch, \
                      float* D, long long d_batch, void* ws, size_t wsb, void* stream)
NSVF_GEMM_DECL(nsvf_gemm_tn_v0); NSVF_GEMM_DECL(nsvf_gemm_tn_v1); NSVF_GEMM_DECL(nsvf_gemm_tn_v2); NSVF_GEMM_DECL(nsvf_gemm_tn_v3); NSVF_GEMM_DECL(nsvf_gemm_tn_v4);
NSVF_GEMM_DECL(nsvf_gemm_nn_v0); NSVF_GEMM_DECL(nsvf_gemm_nn_v1); NSVF_GEMM_DECL(nsvf_gemm_nn_v2); NSVF_GEMM_DECL(nsvf_gemm_nn_v3); NSVF_GEMM_DECL(nsvf_gemm_nn_v4);
NSVF_GEMM_DECL(nsvf_gemm_nt_v0); NSVF_GEMM_DECL(nsvf_gemm_nt_v1); NSVF_GEMM_DECL(nsvf_gemm_nt_v2); NSVF_GEMM_DECL(nsvf_gemm_nt_v3); NSVF_GEMM_DECL(nsvf_gemm_nt_v4);
typedef int (*nsvf_gemm_fn)(int, int, int, int, const float*, long long, const float*, long long, float*, long long, void*,
                            size_t, void*);
static nsvf_gemm_fn g_tn[5] = {nsvf_gemm_tn_v0, nsvf_gemm_tn_v1, nsvf_gemm_tn_v2, nsvf_gemm_tn_v3, nsvf_gemm_tn_v4};
static nsvf_gemm_fn g_nn[5] = {nsvf_gemm_nn_v0, nsvf_gemm_nn_v1, nsvf_gemm_nn_v2, nsvf_gemm_nn_v3, nsvf_gemm_nn_v4};
static nsvf_gemm_fn g_nt[5] = {nsvf_gemm_nt_v0, nsvf_gemm_nt_v1, nsvf_gemm_nt_v2, nsvf_gemm_nt_v3, nsvf_gemm_nt_v4};
static int variant_of(const char* env, int dflt) {
  const char* e = getenv(env);
  const int v = e ? atoi(e) : dflt;
  return v < 0 || v > 4 ? dflt : v;
}
#endif

namespace nsvf {

// dW[c] = sum_l partial[l][c] + sum_{r in [row0, M)} dh[r, n] * x[r, k],  c = n * K + k   (fixed order: deterministic)
__global__ void __launch_bounds__(256)
linear_dw_reduce_kernel(int L, int N, int K, const float* __restrict__ partial, long long row0, long long M,
                        const float* __restrict__ dh, const float* __restrict__ x, float* __restrict__ dW) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N * K) return;
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += partial[(long long)l * N * K + c];
  const int n = c / K, k = c - n * K;
  for (long long r = row0; r < M; ++r) s = fmaf(dh[r * N + n], x[r * K + k], s);
  dW[c] = s;
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_linear_available(void) {
#ifndef NSVF_NO_CUTLASS
  return 1;
#else
  return 0;
#endif
}

extern "C" size_t nsvf_linear_workspace_bytes(long long M, int N, int K, int splits) {
  // per-split partial weight gradients + scratch for the GEMM kernels (they need none today; 1 MiB of headroom)
  return (size_t)splits * (size_t)N * (size_t)K * sizeof(float) + (1u << 20);
}

#ifndef NSVF_NO_CUTLASS
#define NSVF_GEMM_CHECK(call, what)                                     \
  do {                                                                  \
    const int rc_ = (call);                                             \
    if (rc_ != 0) {                                                     \
      nsvf::set_error("%s: CUTLASS status code %d", what, rc_);         \
      return 1;                                                         \
    }                                                                   \
    nsvf::count_launch();                                               \
  } while (0)

extern "C" int nsvf_linear_fwd(nsvf_stream_t stream, long long M, int N, int K, const float* x, const float* W, float* h,
                               void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(M > 0 && M < (1ll << 31) && N % 4 == 0 && K % 4 == 0, "linear_fwd: M > 0, N and K multiples of 4 required");
  NSVF_REQUIRE((((uintptr_t)x | (uintptr_t)W | (uintptr_t)h) & 15) == 0, "linear_fwd: pointers must be 16-byte aligned");
  NSVF_GEMM_CHECK(g_tn[variant_of("NSVF_GEMM_TN", 0)]((int)M, N, K, 1, x, 0, W, 0, h, 0, workspace, workspace_bytes, stream), "linear_fwd");
  return 0;
}

extern "C" int nsvf_linear_bwd_input(nsvf_stream_t stream, long long M, int N, int K, const float* dh, const float* W,
                                     float* dx, void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(M > 0 && M < (1ll << 31) && N % 4 == 0 && K % 4 == 0, "linear_bwd_input: M > 0, N and K multiples of 4 required");
  NSVF_REQUIRE((((uintptr_t)dh | (uintptr_t)W | (uintptr_t)dx) & 15) == 0, "linear_bwd_input: pointers must be 16-byte aligned");
  // dx [M, K] = dh [M, N] * W [N, K]: GEMM (M, K, N)
  NSVF_GEMM_CHECK(g_nn[variant_of("NSVF_GEMM_NN", 0)]((int)M, K, N, 1, dh, 0, W, 0, dx, 0, workspace, workspace_bytes, stream), "linear_bwd_input");
  return 0;
}

extern "C" int nsvf_linear_bwd_weight(nsvf_stream_t stream_, long long M, int N, int K, const float* dh, const float* x,
                                      float* dW, int splits, void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M > 0 && M < (1ll << 31) && N % 4 == 0 && K % 4 == 0 && splits >= 1,
               "linear_bwd_weight: M > 0, N and K multiples of 4, splits >= 1 required");
  NSVF_REQUIRE((((uintptr_t)dh | (uintptr_t)x | (uintptr_t)dW | (uintptr_t)workspace) & 15) == 0,
               "linear_bwd_weight: pointers must be 16-byte aligned");
  const size_t part_bytes = (size_t)splits * N * K * sizeof(float);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= part_bytes, "linear_bwd_weight: workspace too small");
  float* partial = static_cast<float*>(workspace);
  long long blk = (M / splits) & ~3ll;     // rows per split, multiple of 4 (batch strides stay 16-byte aligned)
  int L = splits;
  if (blk < 64) { blk = 0; L = 0; }        // tiny M: everything goes through the tail loop of the reduce kernel
  if (L > 0) {
    // partial[l] [N, K] = dh[l*blk : (l+1)*blk]^T * x[l*blk : (l+1)*blk]: GEMM (N, K, blk) x L
    NSVF_GEMM_CHECK(g_nt[variant_of("NSVF_GEMM_NT", 2)](N, K, (int)blk, L, dh, blk * N, x, blk * K, partial, (long long)N * K,
                                         static_cast<char*>(workspace) + part_bytes, workspace_bytes - part_bytes, stream_),
                    "linear_bwd_weight");
  }
  NSVF_TIMED_LAUNCH("linear_dw_reduce_kernel", stream,
                    (linear_dw_reduce_kernel<<<(N * K + 255) / 256, 256, 0, stream>>>(L, N, K, partial, blk * L, M, dh, x, dW)));
  return 0;
}
#else
extern "C" int nsvf_linear_fwd(nsvf_stream_t, long long, int, int, const float*, const float*, float*, void*, size_t) {
  nsvf::set_error("linear_fwd: built without the CUTLASS headers (NSVF_NO_CUTLASS)");
  return 1;
}
extern "C" int nsvf_linear_bwd_input(nsvf_stream_t, long long, int, int, const float*, const float*, float*, void*, size_t) {
  nsvf::set_error("linear_bwd_input: built without the CUTLASS headers (NSVF_NO_CUTLASS)");
  return 1;
}
extern "C" int nsvf_linear_bwd_weight(nsvf_stream_t, long long, int, int, const float*, const float*, float*, int, void*, size_t) {
  nsvf::set_error("linear_bwd_weight: built without the CUTLASS headers (NSVF_NO_CUTLASS)");
  return 1;
}
#endif
