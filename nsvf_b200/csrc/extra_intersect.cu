// ball_intersect and triangle_intersect for sm_100a — the two clib entry points OUTSIDE the NSVF path (SURVEY.md §8a
// X1: ball_intersect has no caller in the reference, triangle_intersect serves the experimental mesh encoder), so that
// the Level-1 module exposes all 7 functions of fairnr/clib/src/binding.cpp:11-20.
//
// Replaces fairnr/clib/src/intersect_gpu.cu:15-70 (ball) and :240-347 (triangle), which give every ray ONE thread that
// scans all primitives and (triangle) insertion-sorts its hits in place in global memory.  Here a WARP owns a ray:
//   * 32 primitives are tested per step, one per lane (coalesced loads of the primitive arrays, the ray in registers);
//   * hits are appended in ascending primitive index with __ballot_sync + __popc — the order and the truncation at
//     n_max of the reference's sequential scan — and the scan stops as soon as n_max hits are known;
//   * triangle hits are staged in shared memory and ordered by a rank sort on depth (rank = number of hits with a
//     smaller depth); only when two hits of a ray have EQUAL depth does one lane replay the reference's insertion
//     (its carry-forward swap with strict < is not stable: equal depths can end up rotated, and the face order is an
//     integer output), then the cage extents are computed from the sorted neighbours and every output row is
//     written once, coalesced.
// The per-primitive expression trees (distance test, Moeller-Trumbore with __fdividef) are the reference's: they decide
// integer outputs, so nvcc must contract them into the same FMAs (default -fmad=true on both sides).
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

constexpr int kXWarps = 4;

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// Sphere test of intersect_gpu.cu:48-64: hit iff the squared distance of the centre from the ray is < radius^2.
__device__ __forceinline__ bool ball_test(float x0, float y0, float z0, float xw, float yw, float zw, float px, float py,
                                          float pz, float radius2, float& dmin, float& dmax) {
  float x = px - x0;
  float y = py - y0;
  float z = pz - z0;
  float d2 = x * x + y * y + z * z;
  float d2_proj = pow(x * xw + y * yw + z * zw, 2);
  float r2 = d2 - d2_proj;
  if (!(r2 < radius2)) return false;
  float depth = sqrt(d2_proj);
  float depth_blur = sqrt(radius2 - r2);
  dmin = depth - depth_blur;
  dmax = depth + depth_blur;
  return true;
}

__global__ void __launch_bounds__(kXWarps * 32)
ball_intersect_kernel(int b, int n, int m, float radius, int n_max, long long pts_stride,
                      const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                      const float* __restrict__ points_all, int* __restrict__ idx, float* __restrict__ min_depth,
                      float* __restrict__ max_depth) {
  const int lane = threadIdx.x & 31;
  const float radius2 = radius * radius;
  const long long rays = (long long)b * m;
  for (long long j = (long long)blockIdx.x * kXWarps + (threadIdx.x >> 5); j < rays; j += (long long)gridDim.x * kXWarps) {
    const float* points = points_all + (j / m) * pts_stride;
    const float x0 = ray_start[j * 3 + 0], y0 = ray_start[j * 3 + 1], z0 = ray_start[j * 3 + 2];
    const float xw = ray_dir[j * 3 + 0], yw = ray_dir[j * 3 + 1], zw = ray_dir[j * 3 + 2];
    const long long row = j * n_max;
    int cnt = 0;
    for (int k0 = 0; k0 < n && cnt < n_max; k0 += 32) {
      const int k = k0 + lane;
      float dmin = 0.f, dmax = 0.f;
      bool hit = false;
      if (k < n) hit = ball_test(x0, y0, z0, xw, yw, zw, points[k * 3 + 0], points[k * 3 + 1], points[k * 3 + 2], radius2, dmin, dmax);
      const unsigned mask = __ballot_sync(NSVF_FULL_MASK, hit);
      const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && slot < n_max) {
        idx[row + slot] = k;
        min_depth[row + slot] = dmin;
        max_depth[row + slot] = dmax;
      }
      cnt += __popc(mask);
    }
    cnt = cnt < n_max ? cnt : n_max;
    for (int l = cnt + lane; l < n_max; l += 32) {   // the fill the reference gets from -ones / zeros (intersect.cpp:26-34)
      idx[row + l] = -1;
      min_depth[row + l] = 0.f;
      max_depth[row + l] = 0.f;
    }
  }
}

// Moeller-Trumbore of intersect_gpu.cu:240-270 (blurred barycentric bounds, reciprocal by __fdividef); t <= 0 = miss.
__device__ __forceinline__ V3 ray_triangle(V3 ori, V3 dir, V3 v0, V3 v1, V3 v2, float blur) {
  V3 v0v1 = sub3(v1, v0);
  V3 v0v2 = sub3(v2, v0);
  V3 v0O = sub3(ori, v0);
  V3 dir_crs_v0v2 = cross3(dir, v0v2);
  float det = dot3(v0v1, dir_crs_v0v2);
  det = __fdividef(1.0f, det);
  float u = dot3(v0O, dir_crs_v0v2) * det;
  if ((u < 0.0f - blur) || (u > 1.0f + blur)) return v3(-1.0f, 0.0f, 0.0f);
  V3 v0O_crs_v0v1 = cross3(v0O, v0v1);
  float v = dot3(dir, v0O_crs_v0v1) * det;
  if ((v < 0.0f - blur) || (v > 1.0f + blur)) return v3(-1.0f, 0.0f, 0.0f);
  if (((u + v) < 0.0f - blur) || ((u + v) > 1.0f + blur)) return v3(-1.0f, 0.0f, 0.0f);
  float t = dot3(v0v2, v0O_crs_v0v1) * det;
  return v3(t, u, v);
}

__global__ void triangle_intersect_kernel(int b, int n, int m, float cagesize, float blur, int n_max, int warps,
                                          long long face_stride, const float* __restrict__ ray_start,
                                          const float* __restrict__ ray_dir, const float* __restrict__ faces_all,
                                          int* __restrict__ idx, float* __restrict__ depth, float* __restrict__ uv) {
  extern __shared__ float tri_smem[];   // per warp: t[n_max], u[n_max], v[n_max], face[n_max], rank-sorted t[n_max]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* h_t = tri_smem + (size_t)warp * 5 * n_max;
  float* h_u = h_t + n_max;
  float* h_v = h_u + n_max;
  int* h_k = reinterpret_cast<int*>(h_v + n_max);
  float* s_t = h_v + 2 * n_max;
  const long long rays = (long long)b * m;
  for (long long j = (long long)blockIdx.x * warps + warp; j < rays; j += (long long)gridDim.x * warps) {
    const float* faces = faces_all + (j / m) * face_stride;
    const V3 ori = v3(ray_start[j * 3 + 0], ray_start[j * 3 + 1], ray_start[j * 3 + 2]);
    const V3 dir = v3(ray_dir[j * 3 + 0], ray_dir[j * 3 + 1], ray_dir[j * 3 + 2]);
    // 1) the first n_max hit faces in ascending face index
    int cnt = 0;
    for (int k0 = 0; k0 < n && cnt < n_max; k0 += 32) {
      const int k = k0 + lane;
      V3 tuv = v3(-1.0f, 0.0f, 0.0f);
      if (k < n) {
        const float* f = faces + (long long)k * 9;
        tuv = ray_triangle(ori, dir, v3(f[0], f[1], f[2]), v3(f[3], f[4], f[5]), v3(f[6], f[7], f[8]), blur);
      }
      const bool hit = tuv.x > 0;
      const unsigned mask = __ballot_sync(NSVF_FULL_MASK, hit);
      const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && slot < n_max) { h_t[slot] = tuv.x; h_u[slot] = tuv.y; h_v[slot] = tuv.z; h_k[slot] = k; }
      cnt += __popc(mask);
    }
    cnt = cnt < n_max ? cnt : n_max;
    __syncwarp();
    // 2) order by depth.  Distinct depths: rank = number of smaller depths.  Equal depths present: replay the
    //    reference's insertion (:316-323) on the staged hits, in shared memory, by one lane.
    const long long row = j * n_max;
    bool tie = false;
    for (int i = lane; i < cnt; i += 32) {
      const float ti = h_t[i];
      int rank = 0;
      for (int q = 0; q < cnt; ++q) {
        const float tq = h_t[q];
        rank += (tq < ti);
        tie |= (tq == ti) && (q != i);
      }
      s_t[rank] = ti;      // valid only without ties (ranks are then a permutation)
    }
    tie = __any_sync(NSVF_FULL_MASK, tie);
    __syncwarp();
    if (!tie) {
      for (int i = lane; i < cnt; i += 32) {
        const float ti = h_t[i];
        int rank = 0;
        for (int q = 0; q < cnt; ++q) rank += (h_t[q] < ti);
        idx[row + rank] = h_k[i];
        uv[(row + rank) * 2 + 0] = h_u[i];
        uv[(row + rank) * 2 + 1] = h_v[i];
      }
    } else {
      if (lane == 0) {
        // sorted list grows in (s_t, idx row, uv row) exactly like the reference's in-place rows
        for (int c = 0; c < cnt; ++c) {
          int ki = h_k[c];
          float d = h_t[c], u = h_u[c], v = h_v[c];
          for (int l = 0; l < c; ++l) {
            if (d < s_t[l]) {
              const int tk = idx[row + l]; idx[row + l] = ki; ki = tk;
              const float td = s_t[l]; s_t[l] = d; d = td;
              const float tu = uv[(row + l) * 2 + 0]; uv[(row + l) * 2 + 0] = u; u = tu;
              const float tv = uv[(row + l) * 2 + 1]; uv[(row + l) * 2 + 1] = v; v = tv;
            }
          }
          idx[row + c] = ki;
          s_t[c] = d;
          uv[(row + c) * 2 + 0] = u;
          uv[(row + c) * 2 + 1] = v;
        }
      }
    }
    __syncwarp();
    // 3) (t, -cage_near, cage_far) per sorted hit: half the gap to the neighbouring hit, capped at cagesize (:331-345)
    for (int l = lane; l < n_max; l += 32) {
      float t = 0.f, lo = 0.f, hi = 0.f;
      if (l < cnt) {
        t = s_t[l];
        lo = l == 0 ? -cagesize : -fminf(cagesize, .5 * (t - s_t[l - 1]));
        hi = l == cnt - 1 ? cagesize : fminf(cagesize, .5 * (s_t[l + 1] - t));
      } else {
        idx[row + l] = -1;
        uv[(row + l) * 2 + 0] = 0.f;
        uv[(row + l) * 2 + 1] = 0.f;
      }
      depth[(row + l) * 3 + 0] = t;
      depth[(row + l) * 3 + 1] = lo;
      depth[(row + l) * 3 + 2] = hi;
    }
    __syncwarp();
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_ball_intersect(nsvf_stream_t stream_, int b, int n, int m, float radius, int n_max,
                                   const float* ray_start, const float* ray_dir, const float* points,
                                   long long points_batch_stride, int* idx, float* min_depth, float* max_depth) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && n >= 0 && m >= 0 && n_max >= 0, "ball_intersect: negative size");
  if (b == 0 || m == 0 || n_max == 0) return 0;
  const long long rays = (long long)b * m;
  long long want = (rays + kXWarps - 1) / kXWarps, cap = (long long)num_sms() * 16;
  ball_intersect_kernel<<<(int)(want < cap ? want : cap), kXWarps * 32, 0, stream>>>(
      b, n, m, radius, n_max, points_batch_stride, ray_start, ray_dir, points, idx, min_depth, max_depth);
  NSVF_LAUNCH_OK("ball_intersect_kernel");
  return 0;
}

extern "C" int nsvf_triangle_intersect(nsvf_stream_t stream_, int b, int n, int m, float cagesize, float blur,
                                       int n_max, const float* ray_start, const float* ray_dir,
                                       const float* face_points, long long faces_batch_stride, int* idx,
                                       float* depth, float* uv) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && n >= 0 && m >= 0 && n_max >= 0, "triangle_intersect: negative size");
  if (b == 0 || m == 0 || n_max == 0) return 0;
  const size_t per_warp = (size_t)5 * n_max * sizeof(float);
  NSVF_REQUIRE(per_warp <= 200 * 1024, "triangle_intersect: n_max=%d too large for the shared-memory hit list", n_max);
  int warps = (int)((200 * 1024) / per_warp);
  warps = warps > kXWarps ? kXWarps : warps;
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024)
    NSVF_CUDA_OK(cudaFuncSetAttribute(triangle_intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const long long rays = (long long)b * m;
  long long want = (rays + warps - 1) / warps, cap = (long long)num_sms() * 16;
  triangle_intersect_kernel<<<(int)(want < cap ? want : cap), warps * 32, smem, stream>>>(
      b, n, m, cagesize, blur, n_max, warps, faces_batch_stride, ray_start, ray_dir, face_points, idx, depth, uv);
  NSVF_LAUNCH_OK("triangle_intersect_kernel");
  return 0;
}
