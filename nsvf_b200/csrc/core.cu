// Library plumbing: error string, version, device properties.
#include <cstdarg>
#include <cstdio>
#include <cstring>

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

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

static char g_prof_name[64] = "";
static cudaEvent_t g_prof_ev[2] = {nullptr, nullptr};
// accumulating mode (nsvf_profile_begin / nsvf_profile_end): one event pair per launch from a pool owned by the library
constexpr int kProfPool = 8192;
static cudaEvent_t g_pool[kProfPool][2];
static int g_pool_made = 0, g_pool_used = 0, g_pool_dropped = 0;
static bool g_collect = false;
void profile_mark(const char* name, int which, cudaStream_t stream) {
  if (g_prof_name[0] == 0 || strcmp(name, g_prof_name) != 0) return;
  if (g_collect) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
    if (which == 0) {
      if (g_pool_used >= kProfPool) { ++g_pool_dropped; return; }
      if (g_pool_used >= g_pool_made) {
        if (cudaEventCreate(&g_pool[g_pool_made][0]) != cudaSuccess || cudaEventCreate(&g_pool[g_pool_made][1]) != cudaSuccess)
          return;
        ++g_pool_made;
      }
      cudaEventRecord(g_pool[g_pool_used][0], stream);
    } else if (g_pool_used < g_pool_made && g_pool_used < kProfPool && g_pool_dropped == 0) {
      cudaEventRecord(g_pool[g_pool_used][1], stream);
      ++g_pool_used;
    }
    return;
  }
  if (g_prof_ev[which] == nullptr) return;
  // launches that are being captured into a CUDA graph are not timed: recording a caller's event on a capturing
  // stream would tie it to the capture (and invalidate it once the event is used outside)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
  cudaEventRecord(g_prof_ev[which], stream);
}

// reference reciprocal, exported so the CPU oracle can be fed the exact MUFU-based 1/d values
__global__ void ref_rcp_kernel(long long n, const float* __restrict__ x, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = ref_rcp(x[i]);
}

}  // namespace nsvf

extern "C" int nsvf_version(void) { return NSVF_B200_VERSION; }
extern "C" const char* nsvf_last_error(void) { return nsvf::g_err; }

extern "C" unsigned long long nsvf_kernel_launches(void) { return nsvf::g_launches; }

extern "C" int nsvf_profile_kernel(const char* name, void* ev_start, void* ev_stop) {
  if (name == nullptr || name[0] == 0) {
    nsvf::g_prof_name[0] = 0;
    nsvf::g_prof_ev[0] = nsvf::g_prof_ev[1] = nullptr;
    return 0;
  }
  strncpy(nsvf::g_prof_name, name, sizeof(nsvf::g_prof_name) - 1);
  nsvf::g_prof_ev[0] = (cudaEvent_t)ev_start;
  nsvf::g_prof_ev[1] = (cudaEvent_t)ev_stop;
  return 0;
}

extern "C" int nsvf_profile_begin(const char* name) {
  if (name == nullptr || name[0] == 0) return 1;
  strncpy(nsvf::g_prof_name, name, sizeof(nsvf::g_prof_name) - 1);
  nsvf::g_prof_ev[0] = nsvf::g_prof_ev[1] = nullptr;
  nsvf::g_pool_used = 0;
  nsvf::g_pool_dropped = 0;
  nsvf::g_collect = true;
  return 0;
}

extern "C" int nsvf_profile_end(int* n_launches, float* total_ms, float* min_ms, float* max_ms) {
  nsvf::g_collect = false;
  nsvf::g_prof_name[0] = 0;
  float tot = 0.f, mn = 0.f, mx = 0.f;
  int n = 0;
  for (int i = 0; i < nsvf::g_pool_used; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(nsvf::g_pool[i][1]) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&ms, nsvf::g_pool[i][0], nsvf::g_pool[i][1]) != cudaSuccess) continue;
    tot += ms;
    mn = (n == 0 || ms < mn) ? ms : mn;
    mx = ms > mx ? ms : mx;
    ++n;
  }
  if (n_launches) *n_launches = n;
  if (total_ms) *total_ms = tot;
  if (min_ms) *min_ms = mn;
  if (max_ms) *max_ms = mx;
  nsvf::g_pool_used = 0;
  return 0;
}

extern "C" int nsvf_ref_rcp(nsvf_stream_t stream, long long n, const float* x, float* y) {
  if (n <= 0) return 0;
  int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  nsvf::ref_rcp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, x, y);
  NSVF_LAUNCH_OK("ref_rcp_kernel");
  return 0;
}
