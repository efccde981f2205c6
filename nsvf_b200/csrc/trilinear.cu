// Fused trilinear voxel-corner embedding interpolation (gather) and its scatter-add backward, sm_100a.
//
// Replaces the torch-level sequence of SparseVoxelEncoder.forward, fairnr/modules/encoder.py:582-590
// (F.embedding x3 -> [M,8,D] materialisation -> offset_points -> trilinear_interp,
// fairnr/data/geometry.py:195-200) with ONE gather kernel that never materialises the [M,8,D] rows,
// and autograd's embedding backward with one red.global.add.v4.f32 scatter kernel that first merges
// runs of consecutive samples lying in the same voxel (ray-marched samples arrive in ray order).
//
// Math (identical op order to the reference where it is defined):
//   p_a   = (xyz_a - centre[idx]_a) / voxel_size + 0.5                      (IEEE division)
//   w_j   = (q_jx ? p_x : 1 - p_x) * (q_jy ? p_y : 1 - p_y) * (q_jz ? p_z : 1 - p_z),  j = 4*qx + 2*qy + qz
//           (p*q + (1-p)*(1-q) with q in {0,1} reduces to exactly these values)
//   emb   = sum_j w_j * values[feats[idx][j]]
#include <cstdlib>

#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

__device__ __forceinline__ void corner_weights(float px, float py, float pz, float w[8]) {
  const float ax[2] = {__fsub_rn(1.0f, px), px};
  const float ay[2] = {__fsub_rn(1.0f, py), py};
  const float az[2] = {__fsub_rn(1.0f, pz), pz};
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = __fmul_rn(__fmul_rn(ax[(j >> 2) & 1], ay[(j >> 1) & 1]), az[j & 1]);
}

#ifdef NSVF_TRI_EXPERIMENT_NO_RED   // measurement-only build: drop the reductions (results wrong) to bound their cost
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  if (v.x == 123456.789f) *addr = v.y;
}
#else
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
#endif

// D == 32: 8 lanes per sample, one float4 (4 dims) per lane; a warp handles 4 samples per step.
__global__ void __launch_bounds__(256)
trilinear_fwd_d32_kernel(long long M, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                         const int* __restrict__ feats, const float* __restrict__ centres,
                         const float* __restrict__ values, float voxel_size, float* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; s < M;
       s += ((long long)gridDim.x * blockDim.x) >> 3) {
    const int v = sampled_idx[s];
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 0], __ldg(centres + (long long)v * 3 + 0)), voxel_size), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 1], __ldg(centres + (long long)v * 3 + 1)), voxel_size), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 2], __ldg(centres + (long long)v * 3 + 2)), voxel_size), 0.5f);
    float w[8];
    corner_weights(px, py, pz, w);
    const int4 k0 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8));
    const int4 k1 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8) + 1);
    const int key[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float4 e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = __ldg(reinterpret_cast<const float4*>(values + (long long)key[j] * 32) + sub);
    float4 acc = make_float4(w[0] * e[0].x, w[0] * e[0].y, w[0] * e[0].z, w[0] * e[0].w);
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      acc.x = fmaf(w[j], e[j].x, acc.x);
      acc.y = fmaf(w[j], e[j].y, acc.y);
      acc.z = fmaf(w[j], e[j].z, acc.z);
      acc.w = fmaf(w[j], e[j].w, acc.w);
    }
    reinterpret_cast<float4*>(out + s * 32)[sub] = acc;
  }
}

// generic D: one warp per sample, lane strides over the D dims.
__global__ void __launch_bounds__(256)
trilinear_fwd_generic_kernel(long long M, int D, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                             const int* __restrict__ feats, const float* __restrict__ centres,
                             const float* __restrict__ values, float voxel_size, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < M;
       s += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int v = sampled_idx[s];
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 0], centres[(long long)v * 3 + 0]), voxel_size), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 1], centres[(long long)v * 3 + 1]), voxel_size), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 2], centres[(long long)v * 3 + 2]), voxel_size), 0.5f);
    float w[8];
    corner_weights(px, py, pz, w);
    int key[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) key[j] = feats[(long long)v * 8 + j];
    for (int d = lane; d < D; d += 32) {
      float acc = w[0] * values[(long long)key[0] * D + d];
#pragma unroll
      for (int j = 1; j < 8; ++j) acc = fmaf(w[j], values[(long long)key[j] * D + d], acc);
      out[s * D + d] = acc;
    }
  }
}

// Backward, D == 32. Each 8-lane group walks a run of `RUN` consecutive samples and keeps the 8 corner
// gradients (8 x float4 per lane) in registers while the voxel id does not change; one vectorised
// reduction per corner per run segment. grad_xyz (optional) needs the corner rows again:
//   d emb_d / d xyz_a = (1 / voxel_size) * sum_j (d w_j / d p_a) * E_j[d]
constexpr int kTriRun = 8;

__global__ void __launch_bounds__(256)
trilinear_bwd_d32_kernel(long long M, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                         const int* __restrict__ feats, const float* __restrict__ centres,
                         const float* __restrict__ values, float voxel_size, const float* __restrict__ grad_out,
                         float* __restrict__ grad_values, float* __restrict__ grad_xyz) {
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  const long long n_runs = (M + kTriRun - 1) / kTriRun;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n_runs;
       r += ((long long)gridDim.x * blockDim.x) >> 3) {
    const long long s0 = r * kTriRun;
    const long long s1 = min(M, s0 + kTriRun);
    int cur = -1;
    int key[8];
    float4 acc[8];
    for (long long s = s0; s < s1; ++s) {
      const int v = sampled_idx[s];
      if (v != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) red_add_v4(grad_values + (long long)key[j] * 32 + sub * 4, acc[j]);
        }
        cur = v;
        const int4 k0 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8));
        const int4 k1 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8) + 1);
        key[0] = k0.x; key[1] = k0.y; key[2] = k0.z; key[3] = k0.w;
        key[4] = k1.x; key[5] = k1.y; key[6] = k1.z; key[7] = k1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float px = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 0], __ldg(centres + (long long)v * 3 + 0)), voxel_size), 0.5f);
      const float py = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 1], __ldg(centres + (long long)v * 3 + 1)), voxel_size), 0.5f);
      const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 2], __ldg(centres + (long long)v * 3 + 2)), voxel_size), 0.5f);
      float w[8];
      corner_weights(px, py, pz, w);
      const float4 g = __ldg(reinterpret_cast<const float4*>(grad_out + s * 32) + sub);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j].x = fmaf(w[j], g.x, acc[j].x);
        acc[j].y = fmaf(w[j], g.y, acc[j].y);
        acc[j].z = fmaf(w[j], g.z, acc[j].z);
        acc[j].w = fmaf(w[j], g.w, acc[j].w);
      }
      if (grad_xyz != nullptr) {
        const float ax[2] = {1.0f - px, px}, ay[2] = {1.0f - py, py}, az[2] = {1.0f - pz, pz};
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(values + (long long)key[j] * 32) + sub);
          const float dot = g.x * e.x + g.y * e.y + g.z * e.z + g.w * e.w;
          const float sx = ((j >> 2) & 1) ? 1.f : -1.f, sy = ((j >> 1) & 1) ? 1.f : -1.f, sz = (j & 1) ? 1.f : -1.f;
          gx = fmaf(sx * ay[(j >> 1) & 1] * az[j & 1], dot, gx);
          gy = fmaf(sy * ax[(j >> 2) & 1] * az[j & 1], dot, gy);
          gz = fmaf(sz * ax[(j >> 2) & 1] * ay[(j >> 1) & 1], dot, gz);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {  // reduce over the 8 lanes of this sample only
          gx += __shfl_xor_sync(gmask, gx, o);
          gy += __shfl_xor_sync(gmask, gy, o);
          gz += __shfl_xor_sync(gmask, gz, o);
        }
        if (sub < 3) grad_xyz[s * 3 + sub] = (sub == 0 ? gx : (sub == 1 ? gy : gz)) / voxel_size;
      }
    }
    if (cur >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red_add_v4(grad_values + (long long)key[j] * 32 + sub * 4, acc[j]);
    }
  }
}

__global__ void __launch_bounds__(256)
trilinear_bwd_generic_kernel(long long M, int D, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                             const int* __restrict__ feats, const float* __restrict__ centres,
                             const float* __restrict__ values, float voxel_size, const float* __restrict__ grad_out,
                             float* __restrict__ grad_values, float* __restrict__ grad_xyz) {
  const int lane = threadIdx.x & 31;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < M;
       s += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int v = sampled_idx[s];
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 0], centres[(long long)v * 3 + 0]), voxel_size), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 1], centres[(long long)v * 3 + 1]), voxel_size), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(xyz[s * 3 + 2], centres[(long long)v * 3 + 2]), voxel_size), 0.5f);
    float w[8];
    corner_weights(px, py, pz, w);
    const float ax[2] = {1.0f - px, px}, ay[2] = {1.0f - py, py}, az[2] = {1.0f - pz, pz};
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float g = grad_out[s * D + d];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long row = (long long)feats[(long long)v * 8 + j] * D + d;
        atomicAdd(grad_values + row, w[j] * g);
        if (grad_xyz != nullptr) {
          const float dot = g * values[row];
          const float sx = ((j >> 2) & 1) ? 1.f : -1.f, sy = ((j >> 1) & 1) ? 1.f : -1.f, sz = (j & 1) ? 1.f : -1.f;
          gx = fmaf(sx * ay[(j >> 1) & 1] * az[j & 1], dot, gx);
          gy = fmaf(sy * ax[(j >> 2) & 1] * az[j & 1], dot, gy);
          gz = fmaf(sz * ax[(j >> 2) & 1] * ay[(j >> 1) & 1], dot, gz);
        }
      }
    }
    if (grad_xyz != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gx += __shfl_xor_sync(NSVF_FULL_MASK, gx, o);
        gy += __shfl_xor_sync(NSVF_FULL_MASK, gy, o);
        gz += __shfl_xor_sync(NSVF_FULL_MASK, gz, o);
      }
      if (lane < 3) grad_xyz[s * 3 + lane] = (lane == 0 ? gx : (lane == 1 ? gy : gz)) / voxel_size;
    }
  }
}

// ---- D == 32, second generation ---------------------------------------------------------------------------
// A warp takes 32 consecutive samples.  Phase A: one lane per sample computes the 8 corner weights once (the
// first-generation kernel recomputed them in all 8 lanes of a sample's group, IEEE divisions included) and parks
// {weights, corner keys, voxel id} in shared memory.  Phase B: the four 8-lane groups each walk 8 CONSECUTIVE
// samples, one float4 (4 dims) per lane; the 8 corner rows stay in registers while the voxel id does not change
// (ray-marched samples arrive in ray order, ~8 per voxel at step = voxel/8), so most samples cost 4 LDS.128,
// 32 FMAs and one coalesced 128-byte store.
constexpr int kTriRow = 20;   // words per staged sample: 8 weights, 8 keys, voxel id, 3 pad (80 B, 16-B aligned)
// Word offset of staged sample `sl`.  The four 8-lane groups of a warp read rows that are 8 (or 16) samples apart
// in the same LDS instruction; with a plain 20-word stride those rows start at the same bank (8*20 = 160 = 0 mod 32)
// and every broadcast read of {weights, voxel id} is a 4-way bank conflict.  Four extra words per 8 rows move the
// groups to different banks.
__device__ __forceinline__ int tri_row_off(int sl) { return sl * kTriRow + ((sl >> 3) << 2); }
constexpr int tri_stage_words(int samples) { return samples * kTriRow + (samples >> 3) * 4; }

struct TriSample { int v; float x, y, z; };

__device__ __forceinline__ TriSample tri_load(long long s, long long M, const int* __restrict__ sampled_idx,
                                              const float* __restrict__ xyz) {
  TriSample t;
  t.v = -1; t.x = t.y = t.z = 0.f;
  if (s < M) {
    t.v = __ldcs(sampled_idx + s);
    t.x = __ldcs(xyz + s * 3 + 0);
    t.y = __ldcs(xyz + s * 3 + 1);
    t.z = __ldcs(xyz + s * 3 + 2);
  }
  return t;
}

__device__ __forceinline__ void tri_phase_a(const TriSample& t, const int* __restrict__ feats,
                                            const float* __restrict__ centres, float voxel_size, float* row) {
  float w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int4 k0 = make_int4(0, 0, 0, 0), k1 = make_int4(0, 0, 0, 0);
  if (t.v >= 0) {
    const float cx = __ldg(centres + (long long)t.v * 3 + 0), cy = __ldg(centres + (long long)t.v * 3 + 1),
                cz = __ldg(centres + (long long)t.v * 3 + 2);
    k0 = __ldg(reinterpret_cast<const int4*>(feats + (long long)t.v * 8));
    k1 = __ldg(reinterpret_cast<const int4*>(feats + (long long)t.v * 8) + 1);
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(t.x, cx), voxel_size), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(t.y, cy), voxel_size), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(t.z, cz), voxel_size), 0.5f);
    corner_weights(px, py, pz, w);
  }
  reinterpret_cast<float4*>(row)[0] = make_float4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<float4*>(row)[1] = make_float4(w[4], w[5], w[6], w[7]);
  reinterpret_cast<int4*>(row)[2] = k0;
  reinterpret_cast<int4*>(row)[3] = k1;
  reinterpret_cast<int*>(row)[16] = t.v;
}

// SPL = samples per lane in phase A: a warp takes 32*SPL consecutive samples and every 8-lane group walks 8*SPL
// CONSECUTIVE ones, so corner rows (fwd) / corner gradients (bwd) are reused across longer voxel runs.
//
// Run snapping (snap != 0): the boundaries between the four groups' sample ranges are moved forward to the next
// voxel-run start (ballot of "voxel id differs from the previous sample's"), so a run of same-voxel samples is
// never cut between two groups: one corner-row gather (fwd) / one flush of 8 red.v4 per lane (bwd) per run and per
// chunk boundary instead of one per run and per group boundary.
template <int SPL>
__device__ __forceinline__ void tri_group_bounds(const int (&v)[SPL], int lane, int g, int snap, int& r0, int& r1) {
  constexpr int GS = 8 * SPL;   // nominal samples per group
  r0 = g * GS;
  r1 = r0 + GS;
  if (!snap) return;
  unsigned long long starts = 0ull;
#pragma unroll
  for (int q = 0; q < SPL; ++q) {
    int pv = __shfl_up_sync(NSVF_FULL_MASK, v[q], 1);
    if (q > 0) {
      const int last = __shfl_sync(NSVF_FULL_MASK, v[q - 1], 31);
      if (lane == 0) pv = last;
    }
    const unsigned m = __ballot_sync(NSVF_FULL_MASK, (q == 0 && lane == 0) || v[q] != pv);
    starts |= (unsigned long long)m << (32 * q);
  }
  auto bound = [&](int gg) -> int {
    if (gg == 0) return 0;
    if (gg >= 4) return 32 * SPL;
    const unsigned long long mm = starts >> (gg * GS);
    return mm ? gg * GS + __ffsll((long long)mm) - 1 : 32 * SPL;
  };
  r0 = bound(g);
  r1 = bound(g + 1);
}

template <int SPL, int WARPS, bool SNAP>
__global__ void __launch_bounds__(WARPS * 32)
trilinear_fwd_d32_v2_kernel(long long M, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                            const int* __restrict__ feats, const float* __restrict__ centres,
                            const float* __restrict__ values, float voxel_size, float* __restrict__ out) {
  __shared__ __align__(16) float stage[WARPS][tri_stage_words(32 * SPL)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  float* st = stage[warp];
  constexpr int CH = 32 * SPL;
  const long long n_chunks = (M + CH - 1) / CH;
  const long long c_first = (long long)blockIdx.x * WARPS + warp, c_step = (long long)gridDim.x * WARPS;
  TriSample cur[SPL];
#pragma unroll
  for (int q = 0; q < SPL; ++q)
    cur[q] = tri_load(c_first * CH + q * 32 + lane, c_first < n_chunks ? M : 0, sampled_idx, xyz);
  for (long long c = c_first; c < n_chunks; c += c_step) {
    const long long s0 = c * CH;
    int myv[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
      // software pipeline: the next chunk's index / position loads are in flight during this chunk's work
      const TriSample nxt = tri_load((c + c_step) * CH + q * 32 + lane, (c + c_step) < n_chunks ? M : 0, sampled_idx, xyz);
      tri_phase_a(cur[q], feats, centres, voxel_size, st + tri_row_off(q * 32 + lane));
      myv[q] = cur[q].v;
      cur[q] = nxt;
    }
    // compile-time trip count when not snapping: the unrolled loop keeps the next sample's LDS in flight
    int r0 = g * 8 * SPL, r1 = r0 + 8 * SPL;
    if (SNAP) tri_group_bounds<SPL>(myv, lane, g, 1, r0, r1);
    __syncwarp();
    int prev = -2;
    float4 e[8];
#pragma unroll 2
    for (int rr = 0; rr < (SNAP ? r1 - r0 : 8 * SPL); ++rr) {
      const int sl = r0 + rr;
      const float* row = st + tri_row_off(sl);
      const int v = reinterpret_cast<const int*>(row)[16];
      if (v >= 0) {
        if (v != prev) {
          const int4 k0 = reinterpret_cast<const int4*>(row)[2], k1 = reinterpret_cast<const int4*>(row)[3];
          const int key[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j)
            e[j] = __ldg(reinterpret_cast<const float4*>(values + (long long)key[j] * 32) + sub);
          prev = v;
        }
        const float4 w0 = reinterpret_cast<const float4*>(row)[0], w1 = reinterpret_cast<const float4*>(row)[1];
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float4 acc = make_float4(w[0] * e[0].x, w[0] * e[0].y, w[0] * e[0].z, w[0] * e[0].w);
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          acc.x = fmaf(w[j], e[j].x, acc.x);
          acc.y = fmaf(w[j], e[j].y, acc.y);
          acc.z = fmaf(w[j], e[j].z, acc.z);
          acc.w = fmaf(w[j], e[j].w, acc.w);
        }
        __stcs(reinterpret_cast<float4*>(out + (s0 + sl) * 32) + sub, acc);   // streaming store: written once
      }
    }
    __syncwarp();
  }
}

// Backward: same two phases, with the chunk's grad_out rows (32 consecutive samples x 128 B = one contiguous 4 KiB
// block) brought into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier).  NBUF = 2 double-buffers it (the
// next chunk streams in while this one is reduced); NBUF = 1 issues the next chunk's copy right after the last read
// of the buffer, so it lands under the next phase A — half the shared memory, 7 resident CTAs per SM instead of 5,
// which measured 5 % faster.  Each 8-lane group keeps the 8 corner gradients of the current voxel in registers
// across a run of samples and flushes them with red.global.add.v4.f32.
// Measured bound (NSVF_TRI_EXPERIMENT_NO_RED build): 1.14 ms without the reductions vs 1.69 ms with them at 40 M
// samples; the 240 M reduction sectors (7.7 GB, 171 B/sample at 6 samples per voxel run) move over the 64 B/clk
// SM -> L2 port (0.41 ms at 148 SMs), the same port every load request uses, so they do not overlap with the rest.
template <int WARPS, bool SNAP, int NBUF, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
trilinear_bwd_d32_v2_kernel(long long M, const int* __restrict__ sampled_idx, const float* __restrict__ xyz,
                            const int* __restrict__ feats, const float* __restrict__ centres, float voxel_size,
                            const float* __restrict__ grad_out, float* __restrict__ grad_values,
                            unsigned long long perm_mul) {
  __shared__ __align__(16) float stage[WARPS][tri_stage_words(32)];
  __shared__ __align__(128) float gbuf[WARPS][NBUF][32 * 32];
  __shared__ __align__(8) uint64_t bars[WARPS][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  float* st = stage[warp];
  const long long n_chunks = (M + 31) / 32;
  const long long c_first = (long long)blockIdx.x * WARPS + warp, c_step = (long long)gridDim.x * WARPS;
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  // Chunk order: warp w's i-th chunk is perm(c) = c * perm_mul mod n_chunks (perm_mul coprime to n_chunks; 1 = identity).
  // Ray-marched streams are spatially coherent — the ~1000 CTAs in flight would otherwise all work on neighbouring
  // pixels, i.e. on the SAME voxels, and their red.global.add to the same 128-byte rows serialise in L2 (measured on
  // the C3 stream: 3.84 ms in stream order against 1.73 ms for a stream without such coherence).
  auto perm = [&](long long c) -> long long {
    return perm_mul == 1ull ? c : (long long)(((unsigned long long)c * perm_mul) % (unsigned long long)n_chunks);
  };
  auto issue = [&](long long c, int slot) {   // lane 0 only; c is a permuted chunk id
    const long long s0 = c * 32;
    const unsigned bytes = (unsigned)(min((long long)32, M - s0) * 128);
    mbar_expect_tx(&bars[warp][slot], bytes);
    tma_bulk_g2s(gbuf[warp][slot], grad_out + s0 * 32, bytes, &bars[warp][slot]);
  };
  if (lane == 0 && c_first < n_chunks) issue(perm(c_first), 0);
  TriSample cur = tri_load((c_first < n_chunks ? perm(c_first) : 0) * 32 + lane, c_first < n_chunks ? M : 0, sampled_idx, xyz);
  int it = 0;
  for (long long c = c_first; c < n_chunks; c += c_step, ++it) {
    const int slot = NBUF == 2 ? (it & 1) : 0;
    const long long c_nxt = (c + c_step) < n_chunks ? perm(c + c_step) : 0;
    if (NBUF == 2 && lane == 0 && c + c_step < n_chunks) issue(c_nxt, slot ^ 1);   // buffer released by the
    const TriSample nxt = tri_load(c_nxt * 32 + lane, (c + c_step) < n_chunks ? M : 0, sampled_idx, xyz);
    tri_phase_a(cur, feats, centres, voxel_size, st + tri_row_off(lane));      // __syncwarp ending iteration it-1
    const int myv[1] = {cur.v};
    cur = nxt;
    int r0 = g * 8, r1 = r0 + 8;
    if (SNAP) tri_group_bounds<1>(myv, lane, g, 1, r0, r1);
    __syncwarp();
    mbar_wait(&bars[warp][slot], NBUF == 2 ? ((it >> 1) & 1) : (it & 1));
    const float* gb = gbuf[warp][slot];
    int prev = -2;
    int key[8];
    float4 acc[8];
#pragma unroll 2
    for (int rr = 0; rr < (SNAP ? r1 - r0 : 8); ++rr) {
      const int sl = r0 + rr;
      const float* row = st + tri_row_off(sl);
      const int v = reinterpret_cast<const int*>(row)[16];
      if (v >= 0) {
        if (v != prev) {
          if (prev >= 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) red_add_v4(grad_values + (long long)key[j] * 32 + sub * 4, acc[j]);
          }
          const int4 k0 = reinterpret_cast<const int4*>(row)[2], k1 = reinterpret_cast<const int4*>(row)[3];
          key[0] = k0.x; key[1] = k0.y; key[2] = k0.z; key[3] = k0.w;
          key[4] = k1.x; key[5] = k1.y; key[6] = k1.z; key[7] = k1.w;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          prev = v;
        }
        const float4 w0 = reinterpret_cast<const float4*>(row)[0], w1 = reinterpret_cast<const float4*>(row)[1];
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float4 gq = reinterpret_cast<const float4*>(gb + sl * 32)[sub];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j].x = fmaf(w[j], gq.x, acc[j].x);
          acc[j].y = fmaf(w[j], gq.y, acc[j].y);
          acc[j].z = fmaf(w[j], gq.z, acc[j].z);
          acc[j].w = fmaf(w[j], gq.w, acc[j].w);
        }
      }
    }
    if (prev >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red_add_v4(grad_values + (long long)key[j] * 32 + sub * 4, acc[j]);
    }
    __syncwarp();
    // single buffer: every lane has read this chunk's rows; the next chunk streams in under its phase A
    if (NBUF == 1 && lane == 0 && c + c_step < n_chunks) issue(c_nxt, 0);
  }
}

static int g_tri_variant = -1;   // NSVF_TRI_VARIANT env: tuning knob (SPL*10 + warps/4)
static int tri_variant() {
  if (g_tri_variant < 0) {
    const char* e = getenv("NSVF_TRI_VARIANT");
    g_tri_variant = e ? atoi(e) : 0;
  }
  return g_tri_variant;
}

static int tri_snap() {   // NSVF_TRI_SNAP: bit 0 = run snapping in the backward (default on), bit 1 = in the forward
  static int v = getenv("NSVF_TRI_SNAP") ? atoi(getenv("NSVF_TRI_SNAP")) : 1;
  return v;
}

static int grid_for(long long work_items, int items_per_block, int max_blocks_per_sm) {
  long long want = (work_items + items_per_block - 1) / items_per_block;
  long long cap = (long long)num_sms() * max_blocks_per_sm;
  long long g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace nsvf

using namespace nsvf;

extern "C" int nsvf_trilinear_embed_fwd(nsvf_stream_t stream_, long long M, int D, const int* sampled_idx,
                                        const float* sampled_xyz, const int* feats, const float* centres,
                                        const float* values, float voxel_size, float* out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0 && D > 0, "trilinear_embed_fwd: bad sizes");
  if (M == 0) return 0;
  if (D == 32) {
    NSVF_REQUIRE((((uintptr_t)values | (uintptr_t)out | (uintptr_t)feats) & 15) == 0,
                 "trilinear_embed_fwd: values/out/feats must be 16-byte aligned");
    NSVF_REQUIRE(((uintptr_t)values & 15) == 0, "trilinear_embed_fwd: values must be 16-byte aligned");
    static int bps = getenv("NSVF_TRI_BPS") ? atoi(getenv("NSVF_TRI_BPS")) : 24;
    const bool snap = (tri_snap() & 2) != 0;
#define NSVF_TRI_FWD(SPL, W, GRID)                                                                              \
  do {                                                                                                          \
    if (snap) {                                                                                                 \
      NSVF_TIMED_LAUNCH("trilinear_fwd_kernel", stream,                                                         \
                        (trilinear_fwd_d32_v2_kernel<SPL, W, true><<<GRID, W * 32, 0, stream>>>(                \
                            M, sampled_idx, sampled_xyz, feats, centres, values, voxel_size, out)));            \
    } else {                                                                                                    \
      NSVF_TIMED_LAUNCH("trilinear_fwd_kernel", stream,                                                         \
                        (trilinear_fwd_d32_v2_kernel<SPL, W, false><<<GRID, W * 32, 0, stream>>>(               \
                            M, sampled_idx, sampled_xyz, feats, centres, values, voxel_size, out)));            \
    }                                                                                                           \
  } while (0)
    switch (tri_variant() == 0 ? 2 : tri_variant()) {
      case 5:
        NSVF_TRI_FWD(2, 8, grid_for((M + 63) / 64, 8, bps));
        break;
      case 6:
        NSVF_TRI_FWD(2, 2, grid_for((M + 63) / 64, 2, bps));
        break;
      case 2:
        NSVF_TRI_FWD(2, 4, grid_for((M + 63) / 64, 4, bps));
        break;
      case 4:
        NSVF_TRI_FWD(4, 4, grid_for((M + 127) / 128, 4, 12));
        break;
      case 1:
        NSVF_TRI_FWD(1, 4, grid_for((M + 31) / 32, 4, 12));
        break;
      default:
        NSVF_TRI_FWD(1, 8, grid_for((M + 31) / 32, 8, 8));
    }
    return 0;
  } else {
    trilinear_fwd_generic_kernel<<<grid_for(M, 8, 16), 256, 0, stream>>>(M, D, sampled_idx, sampled_xyz, feats,
                                                                         centres, values, voxel_size, out);
  }
  NSVF_LAUNCH_OK("trilinear_fwd_kernel");
  return 0;
}

extern "C" int nsvf_trilinear_embed_bwd(nsvf_stream_t stream_, long long M, int D, const int* sampled_idx,
                                        const float* sampled_xyz, const int* feats, const float* centres,
                                        const float* values, float voxel_size, const float* grad_out,
                                        float* grad_values, float* grad_xyz) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(M >= 0 && D > 0, "trilinear_embed_bwd: bad sizes");
  if (M == 0) return 0;
  if (D == 32) {
    NSVF_REQUIRE((((uintptr_t)values | (uintptr_t)grad_out | (uintptr_t)grad_values | (uintptr_t)feats) & 15) == 0,
                 "trilinear_embed_bwd: values/grad_out/grad_values/feats must be 16-byte aligned");
    if (grad_xyz == nullptr) {
      NSVF_REQUIRE(((uintptr_t)grad_out & 15) == 0, "trilinear_embed_bwd: grad_out must be 16-byte aligned");
      // single grad_out buffer + <= 73 registers: 7 CTAs / SM (1.69 ms vs 1.78 ms for the double-buffered, 5-CTA shape
      // at 40 M samples); NSVF_TRI_BWD=2 selects the double-buffered shape, NSVF_TRI_SNAP=0 the unsnapped groups
      static int bwd_gen = getenv("NSVF_TRI_BWD") ? atoi(getenv("NSVF_TRI_BWD")) : 1;
      // multiplicative permutation of the chunk order (NSVF_TRI_PERM=0 keeps the stream order): a prime that does not
      // divide the chunk count is coprime to it
      static int use_perm = getenv("NSVF_TRI_PERM") ? atoi(getenv("NSVF_TRI_PERM")) : 1;
      const unsigned long long n_chunks = (unsigned long long)((M + 31) / 32);
      unsigned long long perm_mul = 1ull;
      if (use_perm && n_chunks > 4096) {
        const unsigned long long primes[3] = {1000003ull, 999983ull, 1299709ull};
        for (int i = 0; i < 3 && perm_mul == 1ull; ++i)
          if (n_chunks % primes[i] != 0) perm_mul = primes[i];
      }
      if (bwd_gen == 2 && (tri_snap() & 1)) {
        NSVF_TIMED_LAUNCH("trilinear_bwd_kernel", stream,
                          (trilinear_bwd_d32_v2_kernel<4, true, 2, 5><<<grid_for((M + 31) / 32, 4, 5), 128, 0, stream>>>(
                              M, sampled_idx, sampled_xyz, feats, centres, voxel_size, grad_out, grad_values, perm_mul)));
      } else if (bwd_gen == 2) {
        NSVF_TIMED_LAUNCH("trilinear_bwd_kernel", stream,
                          (trilinear_bwd_d32_v2_kernel<4, false, 2, 5><<<grid_for((M + 31) / 32, 4, 5), 128, 0, stream>>>(
                              M, sampled_idx, sampled_xyz, feats, centres, voxel_size, grad_out, grad_values, perm_mul)));
      } else {
        NSVF_TIMED_LAUNCH("trilinear_bwd_kernel", stream,
                          (trilinear_bwd_d32_v2_kernel<4, true, 1, 7><<<grid_for((M + 31) / 32, 4, 7), 128, 0, stream>>>(
                              M, sampled_idx, sampled_xyz, feats, centres, voxel_size, grad_out, grad_values, perm_mul)));
      }
      return 0;
    }
    const long long runs = (M + kTriRun - 1) / kTriRun;
    trilinear_bwd_d32_kernel<<<grid_for(runs, 32, 16), 256, 0, stream>>>(M, sampled_idx, sampled_xyz, feats, centres,
                                                                         values, voxel_size, grad_out, grad_values,
                                                                         grad_xyz);
  } else {
    trilinear_bwd_generic_kernel<<<grid_for(M, 8, 16), 256, 0, stream>>>(M, D, sampled_idx, sampled_xyz, feats,
                                                                         centres, values, voxel_size, grad_out,
                                                                         grad_values, grad_xyz);
  }
  NSVF_LAUNCH_OK("trilinear_bwd_kernel");
  return 0;
}
