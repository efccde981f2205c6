// Shared device/host helpers for the nsvf_b200 kernels (sm_100a only).
//
// Arithmetic contract (see DESIGN.md "Arithmetic contract"): every floating-point operation that
// feeds an integer decision of the reference kernels is written with explicit round-to-nearest
// intrinsics (__fadd_rn/__fsub_rn/__fmul_rn/__fmaf_rn/__fdiv_rn), which nvcc never contracts or
// re-associates, so the SASS op sequence equals the reference's
// (fairnr/clib/src/intersect_gpu.cu:73-122, fairnr/clib/src/sample_gpu.cu:15-202 compiled with
// default -fmad=true -prec-div=true).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NSVF_FULL_MASK 0xffffffffu

namespace nsvf {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define NSVF_CUDA_OK(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return nsvf::cuda_fail(_e, #expr); \
  } while (0)
#define NSVF_LAUNCH_OK(name)                                       \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return nsvf::cuda_fail(_e, name);        \
    nsvf::count_launch();                                          \
  } while (0)
// Launch `stmt` (a <<<>>> expression on `stream`), bracketed by the profiling events registered for `name`
// through nsvf_profile_kernel (no-op when none are registered).
#define NSVF_TIMED_LAUNCH(name, stream, stmt)                       \
  do {                                                             \
    nsvf::profile_mark(name, 0, stream);                           \
    stmt;                                                          \
    nsvf::profile_mark(name, 1, stream);                           \
    NSVF_LAUNCH_OK(name);                                          \
  } while (0)
#define NSVF_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      nsvf::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

int num_sms();  // SM count of the current device (148 on B200), cached
void count_launch();
void profile_mark(const char* name, int which, cudaStream_t stream);

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// The reference's reciprocal: __fdividef(1.0f, d)  (intersect_gpu.cu:86-90).
__device__ __forceinline__ float ref_rcp(float d) { return __fdividef(1.0f, d); }

// True when the reference slab test can never see a NaN for this ray component: finite origin and a
// finite, non-zero reciprocal (finite * finite is never NaN). Voxel centres are assumed finite.
__device__ __forceinline__ bool regular_component(float o, float inv) {
  return (fabsf(o) <= 3.0e38f) && (fabsf(inv) <= 3.0e38f) && (inv != 0.0f);
}

// Reference-exact slab test, select-based (NaN behaviour identical to the reference's
// if/else chain, intersect_gpu.cu:93-120). lo/hi are the box bounds (c - hv, c + hv).
// Returns hit; (tn, tf) = (f_low, f_high).
__device__ __forceinline__ bool slab_exact(float ox, float oy, float oz, float ix, float iy, float iz,
                                           float lx, float ly, float lz, float hx, float hy, float hz,
                                           float& tn, float& tf) {
  float f_low = 0.0f, f_high = 100000.0f;
  float a, b, t;
#define NSVF_AXIS(l, h, o, inv)                         \
  a = __fmul_rn(__fsub_rn(l, o), inv);                  \
  b = __fmul_rn(__fsub_rn(h, o), inv);                  \
  if (b < a) { t = a; a = b; b = t; }                   \
  if (b < f_low) return false;                          \
  if (a > f_high) return false;                         \
  f_low = (a > f_low) ? a : f_low;                      \
  f_high = (b < f_high) ? b : f_high;                   \
  if (f_low > f_high) return false;
  NSVF_AXIS(lx, hx, ox, ix)
  NSVF_AXIS(ly, hy, oy, iy)
  NSVF_AXIS(lz, hz, oz, iz)
#undef NSVF_AXIS
  tn = f_low;
  tf = f_high;
  return true;
}

// Fast slab test, valid (bit-identical hit decision and depths, up to the sign of a zero) when all
// three components are `regular`: (near, far) are the box bounds pre-selected by the sign of inv so
// that near <= far after the multiply (rounding is monotone) and no swap is needed.
__device__ __forceinline__ bool slab_sorted(float ox, float oy, float oz, float ix, float iy, float iz,
                                            float nx, float ny, float nz, float fx, float fy, float fz,
                                            float& tn, float& tf) {
  float t0 = __fmul_rn(__fsub_rn(nx, ox), ix);
  float t1 = __fmul_rn(__fsub_rn(fx, ox), ix);
  float t2 = __fmul_rn(__fsub_rn(ny, oy), iy);
  float t3 = __fmul_rn(__fsub_rn(fy, oy), iy);
  float t4 = __fmul_rn(__fsub_rn(nz, oz), iz);
  float t5 = __fmul_rn(__fsub_rn(fz, oz), iz);
  tn = fmaxf(fmaxf(fmaxf(0.0f, t0), t2), t4);
  tf = fminf(fminf(fminf(100000.0f, t1), t3), t5);
  return tn <= tf;
}

// Conservative slab test for ENCLOSING boxes (internal nodes of our own hierarchy). Uses the same
// (b - o) * inv op sequence (monotone rounding => a box that encloses a voxel box can only widen the
// interval) with fmin/fmax ordering; NaNs (0 * inf) drop out of fmin/fmax, which is safe because the
// node boxes are STRICTLY larger than every box they enclose.
__device__ __forceinline__ bool slab_enclosing(float ox, float oy, float oz, float ix, float iy, float iz,
                                               float lx, float ly, float lz, float hx, float hy, float hz) {
  float a0 = __fmul_rn(__fsub_rn(lx, ox), ix), b0 = __fmul_rn(__fsub_rn(hx, ox), ix);
  float a1 = __fmul_rn(__fsub_rn(ly, oy), iy), b1 = __fmul_rn(__fsub_rn(hy, oy), iy);
  float a2 = __fmul_rn(__fsub_rn(lz, oz), iz), b2 = __fmul_rn(__fsub_rn(hz, oz), iz);
  float tn = fmaxf(fmaxf(fmaxf(0.0f, fminf(a0, b0)), fminf(a1, b1)), fminf(a2, b2));
  float tf = fminf(fminf(fminf(100000.0f, fmaxf(a0, b0)), fmaxf(a1, b1)), fmaxf(a2, b2));
  return !(tn > tf);
}

// Order-preserving integer image of a float (signed compare of the keys == compare of the floats, -0 < +0): lets
// atomicMin / __reduce_max_sync work on floats.
__device__ __forceinline__ int float_order_key(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float float_from_order_key(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NSVF_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NSVF_DONE_%=;\n"
      "bra NSVF_WAIT_%=;\n"
      "NSVF_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// dst (shared), src (global) 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

#endif  // __CUDACC__

}  // namespace nsvf
