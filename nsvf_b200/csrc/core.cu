// Library plumbing: error string, version, device properties.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 1000 + (int)e;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// reference reciprocal, exported so the CPU oracle can be fed the exact MUFU-based 1/d values
__global__ void ref_rcp_kernel(long long n, const float* __restrict__ x, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = ref_rcp(x[i]);
}

}  // namespace nsvf

extern "C" int nsvf_version(void) { return NSVF_B200_VERSION; }
extern "C" const char* nsvf_last_error(void) { return nsvf::g_err; }

extern "C" int nsvf_ref_rcp(nsvf_stream_t stream, long long n, const float* x, float* y) {
  if (n <= 0) return 0;
  int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  nsvf::ref_rcp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, x, y);
  NSVF_LAUNCH_OK("ref_rcp_kernel");
  return 0;
}
