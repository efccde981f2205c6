// ball_intersect and triangle_intersect — the two clib entry points OUTSIDE the NSVF path (SURVEY.md §8a X1:
// ball_intersect has no caller in the reference, triangle_intersect serves the experimental mesh encoder).  They
// exist so that the Level-1 module exposes all 7 functions of fairnr/clib/src/binding.cpp:11-20.  Restated from
// fairnr/clib/src/intersect_gpu.cu:15-70 (ball) and :240-347 (triangle): one thread per ray over all primitives,
// same expression trees (nvcc contracts them into the same FMAs), same in-place insertion sort; the kernels
// themselves write the -1 / 0 fill the reference gets from torch::zeros, and read each ray once.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

__global__ void ball_intersect_kernel(int n, int m, float radius, int n_max, long long pts_stride,
                                      const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                                      const float* __restrict__ points_all, int* __restrict__ idx,
                                      float* __restrict__ min_depth, float* __restrict__ max_depth) {
  const float* points = points_all + (long long)blockIdx.y * pts_stride;
  const float radius2 = radius * radius;
  for (int jj = blockIdx.x * blockDim.x + threadIdx.x; jj < m; jj += gridDim.x * blockDim.x) {
    const long long j = (long long)blockIdx.y * m + jj;
    const float x0 = ray_start[j * 3 + 0], y0 = ray_start[j * 3 + 1], z0 = ray_start[j * 3 + 2];
    const float xw = ray_dir[j * 3 + 0], yw = ray_dir[j * 3 + 1], zw = ray_dir[j * 3 + 2];
    for (int l = 0; l < n_max; ++l) { idx[j * n_max + l] = -1; min_depth[j * n_max + l] = 0.f; max_depth[j * n_max + l] = 0.f; }
    for (int k = 0, cnt = 0; k < n && cnt < n_max; ++k) {
      float x = points[k * 3 + 0] - x0;
      float y = points[k * 3 + 1] - y0;
      float z = points[k * 3 + 2] - z0;
      float d2 = x * x + y * y + z * z;
      float d2_proj = pow(x * xw + y * yw + z * zw, 2);
      float r2 = d2 - d2_proj;
      if (r2 < radius2) {
        idx[j * n_max + cnt] = k;
        float depth = sqrt(d2_proj);
        float depth_blur = sqrt(radius2 - r2);
        min_depth[j * n_max + cnt] = depth - depth_blur;
        max_depth[j * n_max + cnt] = depth + depth_blur;
        ++cnt;
      }
    }
  }
}

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

__global__ void triangle_intersect_kernel(int n, int m, float cagesize, float blur, int n_max, long long face_stride,
                                          const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                                          const float* __restrict__ faces_all, int* __restrict__ idx,
                                          float* __restrict__ depth, float* __restrict__ uv) {
  const float* face_points = faces_all + (long long)blockIdx.y * face_stride;
  for (int jj = blockIdx.x * blockDim.x + threadIdx.x; jj < m; jj += gridDim.x * blockDim.x) {
    const long long j = (long long)blockIdx.y * m + jj;
    const V3 ori = v3(ray_start[j * 3 + 0], ray_start[j * 3 + 1], ray_start[j * 3 + 2]);
    const V3 dir = v3(ray_dir[j * 3 + 0], ray_dir[j * 3 + 1], ray_dir[j * 3 + 2]);
    for (int l = 0; l < n_max; ++l) idx[j * n_max + l] = -1;
    for (int l = 0; l < n_max * 3; ++l) depth[j * n_max * 3 + l] = 0.f;
    for (int l = 0; l < n_max * 2; ++l) uv[j * n_max * 2 + l] = 0.f;
    int cnt = 0;
    for (int k = 0; k < n && cnt < n_max; ++k) {
      const float* f = face_points + (long long)k * 9;
      V3 tuv = ray_triangle(ori, dir, v3(f[0], f[1], f[2]), v3(f[3], f[4], f[5]), v3(f[6], f[7], f[8]), blur);
      if (tuv.x > 0) {
        int ki = k;
        float d = tuv.x, u = tuv.y, v = tuv.z;
        for (int l = 0; l < cnt; l++) {   // insertion by depth (reference :316-323)
          if (d < depth[j * n_max * 3 + l * 3]) {
            int ti = idx[j * n_max + l]; idx[j * n_max + l] = ki; ki = ti;
            float td = depth[j * n_max * 3 + l * 3]; depth[j * n_max * 3 + l * 3] = d; d = td;
            float tu = uv[j * n_max * 2 + l * 2]; uv[j * n_max * 2 + l * 2] = u; u = tu;
            float tv = uv[j * n_max * 2 + l * 2 + 1]; uv[j * n_max * 2 + l * 2 + 1] = v; v = tv;
          }
        }
        idx[j * n_max + cnt] = ki;
        depth[j * n_max * 3 + cnt * 3] = d;
        uv[j * n_max * 2 + cnt * 2] = u;
        uv[j * n_max * 2 + cnt * 2 + 1] = v;
        cnt++;
      }
    }
    for (int l = 0; l < cnt; l++) {   // cage extents between neighbouring hits (reference :331-345)
      if (l == 0) depth[j * n_max * 3 + l * 3 + 1] = -cagesize;
      else depth[j * n_max * 3 + l * 3 + 1] =
               -fminf(cagesize, .5 * (depth[j * n_max * 3 + l * 3] - depth[j * n_max * 3 + l * 3 - 3]));
      if (l == cnt - 1) depth[j * n_max * 3 + l * 3 + 2] = cagesize;
      else depth[j * n_max * 3 + l * 3 + 2] =
               fminf(cagesize, .5 * (depth[j * n_max * 3 + l * 3 + 3] - depth[j * n_max * 3 + l * 3]));
    }
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
  dim3 grid((m + 127) / 128, b);
  ball_intersect_kernel<<<grid, 128, 0, stream>>>(n, m, radius, n_max, points_batch_stride, ray_start, ray_dir, points,
                                                  idx, min_depth, max_depth);
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
  dim3 grid((m + 127) / 128, b);
  triangle_intersect_kernel<<<grid, 128, 0, stream>>>(n, m, cagesize, blur, n_max, faces_batch_stride, ray_start,
                                                      ray_dir, face_points, idx, depth, uv);
  NSVF_LAUNCH_OK("triangle_intersect_kernel");
  return 0;
}
