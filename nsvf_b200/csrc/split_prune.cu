// Half-voxel splitting and voxel pruning for sm_100a.
//
// SPLIT — replaces splitting_points, fairnr/data/geometry.py:250-274 (with discretize_points :241-247 and
// offset_points :229-238), which materialises several [64n,3] int64 tensors and runs torch.unique(dim=0) (a
// sort of 64n rows) plus a scatter_ with duplicate indices whose winner is undefined on CUDA.  Here:
//   * every voxel contributes 8 children x 8 corners = 64 candidate corner keys on the integer lattice
//     (quarter-voxel units); a key is identified by its lexicographic rank, obtained WITHOUT sorting from a
//     dense occupancy bitmap of the lattice: mark (atomicOr) -> per-word popcount prefix scan -> rank =
//     prefix[word] + popc(bits below).  x-major linearisation == torch.unique's lexicographic row order.
//   * the parent of a new key is the MINIMUM voxel index touching it (atomicMin): deterministic.
//   * new embeddings = the parent's trilinear interpolant at p = (key - parent_coord)/4 + 1/2.
//
// PRUNE — replaces SparseVoxelEncoder.get_scores / pruning, fairnr/modules/encoder.py:605-654:
//   * nsvf_prune_lattice_embed: the bits^3 lattice points of each voxel (offset_points(..., bits=16)) are
//     generated and interpolated in one pass; the 8 corner rows of a voxel are loaded once and reused for all
//     of its lattice points (the reference gathers 8 rows per point);
//   * nsvf_prune_keep: keep = (1 - min_l exp(-relu(sigma_l))) > th, one warp per voxel.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

struct SplitGrid {
  float pmin[3];     // per-axis minimum of the voxel centres
  float quarter;     // quarter voxel
  int dim[3];        // lattice extent in keys (max coord + 5)
};

__device__ __forceinline__ void split_old_coord(const SplitGrid& g, const float* __restrict__ points, int v, int c[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a)  // discretize_points: ((p - min) / voxel_size).round_()  (round half to even)
    c[a] = (int)rintf(__fdiv_rn(__fsub_rn(points[(long long)v * 3 + a], g.pmin[a]), g.quarter));
}
__device__ __forceinline__ long long split_lin(const SplitGrid& g, int kx, int ky, int kz) {
  return ((long long)(kx + 2) * g.dim[1] + (ky + 2)) * g.dim[2] + (kz + 2);
}

// one thread per (voxel, child, corner): mark the key; the corner-0 thread also writes the child centre
__global__ void split_mark_kernel(SplitGrid g, int n, const float* __restrict__ points, unsigned* __restrict__ bitmap,
                                  float* __restrict__ new_points) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * 64) return;
  const int v = (int)(t >> 6), child = (int)(t >> 3) & 7, corner = (int)t & 7;
  int c[3];
  split_old_coord(g, points, v, c);
  const int sx = (child >> 2) & 1, sy = (child >> 1) & 1, sz = child & 1;
  const int qx = (corner >> 2) & 1, qy = (corner >> 1) & 1, qz = corner & 1;
  const int kx = c[0] + (2 * sx - 1) + (2 * qx - 1), ky = c[1] + (2 * sy - 1) + (2 * qy - 1),
            kz = c[2] + (2 * sz - 1) + (2 * qz - 1);
  const long long lin = split_lin(g, kx, ky, kz);
  atomicOr(bitmap + (lin >> 5), 1u << (lin & 31));
  if (corner == 0 && new_points != nullptr) {   // offset_points(point_xyz, quarter): c + (+-1) * quarter
    const long long o = ((long long)v * 8 + child) * 3;
    new_points[o + 0] = __fadd_rn(points[(long long)v * 3 + 0], __fmul_rn((float)(2 * sx - 1), g.quarter));
    new_points[o + 1] = __fadd_rn(points[(long long)v * 3 + 1], __fmul_rn((float)(2 * sy - 1), g.quarter));
    new_points[o + 2] = __fadd_rn(points[(long long)v * 3 + 2], __fmul_rn((float)(2 * sz - 1), g.quarter));
  }
}

// ---- popcount prefix scan over the bitmap words (3 phases, any size) ---------------------------------------
constexpr int kScanThreads = 1024, kScanPerThread = 4, kScanBlock = kScanThreads * kScanPerThread;

__device__ __forceinline__ int block_excl_scan(int x, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(NSVF_FULL_MASK, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(NSVF_FULL_MASK, wi, o);
      if (lane >= o) wi += y;
    }
    warp_sums[lane] = wi - w;   // exclusive
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  const int r = incl - x + warp_sums[warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(long long n_words, const unsigned* __restrict__ bitmap, int* __restrict__ block_sums) {
  __shared__ int total;
  const long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanPerThread;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i)
    if (base + i < n_words) s += __popc(bitmap[base + i]);
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_offsets_kernel(int n_blocks, int* __restrict__ block_sums, int* __restrict__ grand_total) {
  __shared__ int total;
  int carry = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
    const int i = b0 + threadIdx.x;
    const int x = i < n_blocks ? block_sums[i] : 0;
    const int e = block_excl_scan(x, &total);
    if (i < n_blocks) block_sums[i] = carry + e;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_words_kernel(long long n_words, const unsigned* __restrict__ bitmap, const int* __restrict__ block_sums,
                  int* __restrict__ word_rank) {
  __shared__ int total;
  const long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanPerThread;
  int c[kScanPerThread], s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    c[i] = base + i < n_words ? __popc(bitmap[base + i]) : 0;
    s += c[i];
  }
  int e = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    if (base + i < n_words) word_rank[base + i] = e;
    e += c[i];
  }
}

__device__ __forceinline__ int split_rank(const unsigned* __restrict__ bitmap, const int* __restrict__ word_rank,
                                          long long lin) {
  const long long w = lin >> 5;
  return word_rank[w] + __popc(bitmap[w] & ((1u << (lin & 31)) - 1u));
}

// one thread per (voxel, child, corner): new_feats + parent = min voxel index
__global__ void split_feats_kernel(SplitGrid g, int n, const float* __restrict__ points,
                                   const unsigned* __restrict__ bitmap, const int* __restrict__ word_rank,
                                   int* __restrict__ new_feats, int* __restrict__ parent) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * 64) return;
  const int v = (int)(t >> 6), child = (int)(t >> 3) & 7, corner = (int)t & 7;
  int c[3];
  split_old_coord(g, points, v, c);
  const int kx = c[0] + (2 * ((child >> 2) & 1) - 1) + (2 * ((corner >> 2) & 1) - 1);
  const int ky = c[1] + (2 * ((child >> 1) & 1) - 1) + (2 * ((corner >> 1) & 1) - 1);
  const int kz = c[2] + (2 * (child & 1) - 1) + (2 * (corner & 1) - 1);
  const int rank = split_rank(bitmap, word_rank, split_lin(g, kx, ky, kz));
  new_feats[t] = rank;
  atomicMin(parent + rank, v);
}

// one warp per voxel: for each of its 27 lattice keys that it owns (parent == voxel) write the key and the
// interpolated embedding.  The voxel's 8 corner rows are read once.
__global__ void __launch_bounds__(256)
split_values_kernel(SplitGrid g, int n, int D, const float* __restrict__ points, const int* __restrict__ feats,
                    const float* __restrict__ values, const unsigned* __restrict__ bitmap,
                    const int* __restrict__ word_rank, const int* __restrict__ parent, int* __restrict__ new_keys,
                    float* __restrict__ new_values) {
  const int lane = threadIdx.x & 31;
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= n) return;
  int c[3];
  split_old_coord(g, points, v, c);
  int key[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) key[j] = feats[(long long)v * 8 + j];
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = d < D ? values[(long long)key[j] * D + d] : 0.f;
    for (int k = 0; k < 27; ++k) {
      const int dx = (k / 9) * 2 - 2, dy = ((k / 3) % 3) * 2 - 2, dz = (k % 3) * 2 - 2;
      const int rank = split_rank(bitmap, word_rank, split_lin(g, c[0] + dx, c[1] + dy, c[2] + dz));
      if (parent[rank] != v) continue;   // warp-uniform
      // p = (key - old_coord) * .25 + .5 in {0, .5, 1};  w_j = prod_a (p*q + (1-p)*(1-q))
      const float px = (float)dx * .25f + 0.5f, py = (float)dy * .25f + 0.5f, pz = (float)dz * .25f + 0.5f;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float wx = ((j >> 2) & 1) ? px : 1.0f - px, wy = ((j >> 1) & 1) ? py : 1.0f - py,
                    wz = (j & 1) ? pz : 1.0f - pz;
        acc += __fmul_rn(__fmul_rn(wx, wy), wz) * e[j];
      }
      if (d < D) new_values[(long long)rank * D + d] = acc;
      if (d0 == 0 && lane < 3 && new_keys != nullptr)
        new_keys[(long long)rank * 3 + lane] = lane == 0 ? c[0] + dx : (lane == 1 ? c[1] + dy : c[2] + dz);
    }
  }
}

// ---- pruning -------------------------------------------------------------------------------------------------
// emb of the bits^3 lattice points of voxels [v0, v0 + nv).  D == 32: 8 lanes x float4 per voxel.
__global__ void __launch_bounds__(256)
prune_lattice_embed_d32_kernel(int nv, int v0, int bits, const int* __restrict__ feats,
                               const float* __restrict__ centres, const float* __restrict__ values, float voxel_size,
                               float half_voxel, float* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const int L = bits * bits * bits;
  // grid: blockIdx.y = voxel (relative), the block's 32 lane-groups stride over the lattice points
  const int v = v0 + blockIdx.y;
  const float cx = centres[(long long)v * 3 + 0], cy = centres[(long long)v * 3 + 1], cz = centres[(long long)v * 3 + 2];
  const int4 k0 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8));
  const int4 k1 = __ldg(reinterpret_cast<const int4*>(feats + (long long)v * 8) + 1);
  const int key[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  float4 e[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) e[j] = __ldg(reinterpret_cast<const float4*>(values + (long long)key[j] * 32) + sub);
  const float inv_b = (float)(bits - 1);
  for (int l = blockIdx.x * 32 + (threadIdx.x >> 3); l < L; l += gridDim.x * 32) {
    const int ix = l / (bits * bits), iy = (l / bits) % bits, iz = l % bits;
    // offset_points(bits): (c - bits) / (bits - 1) with c = 1,3,...,2*bits-1, then point + offset * half_voxel
    const float ox = __fdiv_rn((float)(2 * ix + 1 - bits), inv_b), oy = __fdiv_rn((float)(2 * iy + 1 - bits), inv_b),
                oz = __fdiv_rn((float)(2 * iz + 1 - bits), inv_b);
    const float x = __fadd_rn(cx, __fmul_rn(ox, half_voxel)), y = __fadd_rn(cy, __fmul_rn(oy, half_voxel)),
                z = __fadd_rn(cz, __fmul_rn(oz, half_voxel));
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(x, cx), voxel_size), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(y, cy), voxel_size), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(z, cz), voxel_size), 0.5f);
    const float ax[2] = {__fsub_rn(1.0f, px), px}, ay[2] = {__fsub_rn(1.0f, py), py}, az[2] = {__fsub_rn(1.0f, pz), pz};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float w = __fmul_rn(__fmul_rn(ax[(j >> 2) & 1], ay[(j >> 1) & 1]), az[j & 1]);
      acc.x = fmaf(w, e[j].x, acc.x); acc.y = fmaf(w, e[j].y, acc.y);
      acc.z = fmaf(w, e[j].z, acc.z); acc.w = fmaf(w, e[j].w, acc.w);
    }
    reinterpret_cast<float4*>(out + ((long long)blockIdx.y * L + l) * 32)[sub] = acc;
  }
}

// keep[v] = (1 - min_l exp(-relu(sigma[v, l]))) > th ; min_score optional
__global__ void __launch_bounds__(256)
prune_keep_kernel(int nv, int L, const float* __restrict__ sigma, float th, unsigned char* __restrict__ keep,
                  float* __restrict__ min_score) {
  const int lane = threadIdx.x & 31;
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= nv) return;
  float m = -INFINITY;
  for (int l = lane; l < L; l += 32) m = fmaxf(m, sigma[(long long)v * L + l]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(NSVF_FULL_MASK, m, o));
  if (lane == 0) {
    // exp(-relu(.)) is non-increasing, so min_l exp(-relu(sigma_l)) == exp(-relu(max_l sigma_l)), bit for bit
    const float s = expf(-fmaxf(m, 0.0f));
    if (min_score != nullptr) min_score[v] = s;
    keep[v] = (1.0f - s) > th;
  }
}

}  // namespace nsvf

using namespace nsvf;

static int split_grid(const float* pmin, const int* max_coord, float half_voxel, SplitGrid& g, long long& n_words) {
  g.quarter = half_voxel * 0.5f;
  long long bits = 1;
  for (int a = 0; a < 3; ++a) {
    g.pmin[a] = pmin[a];
    NSVF_REQUIRE(max_coord[a] >= 0, "split: negative lattice extent");
    g.dim[a] = max_coord[a] + 5;
    bits *= g.dim[a];
  }
  NSVF_REQUIRE(bits < (1ll << 38), "split: key lattice of %lld cells is too large for the bitmap dedup", bits);
  n_words = (bits + 31) / 32;
  return 0;
}

extern "C" size_t nsvf_split_workspace_bytes(const int* max_coord) {
  long long bits = 1;
  for (int a = 0; a < 3; ++a) bits *= (long long)(max_coord[a] + 5);
  const long long n_words = (bits + 31) / 32;
  const long long n_blocks = (n_words + kScanBlock - 1) / kScanBlock;
  // bitmap + word_rank + block sums + grand total (128-byte aligned sections)
  return (size_t)(((n_words * 4 + 127) / 128 * 128) * 2 + (n_blocks * 4 + 127) / 128 * 128 + 128);
}

// Phase 1: mark + scan.  pmin f32[3] / max_coord i32[3] are HOST values (per-axis min of points and max of
// round((p - min) / quarter)).  Writes new_points [8n,3] and the number of unique keys to *n_keys (device i32).
extern "C" int nsvf_split_mark(nsvf_stream_t stream_, int n, const float* points, float half_voxel, const float* pmin,
                               const int* max_coord, float* new_points, int* n_keys, void* workspace,
                               size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(n > 0, "split: empty voxel set");
  SplitGrid g;
  long long n_words;
  if (split_grid(pmin, max_coord, half_voxel, g, n_words)) return 1;
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= nsvf_split_workspace_bytes(max_coord),
               "split: workspace too small");
  const long long sec = (n_words * 4 + 127) / 128 * 128;
  unsigned* bitmap = (unsigned*)workspace;
  int* word_rank = (int*)((char*)workspace + sec);
  int* block_sums = (int*)((char*)workspace + 2 * sec);
  const int n_blocks = (int)((n_words + kScanBlock - 1) / kScanBlock);
  NSVF_CUDA_OK(cudaMemsetAsync(bitmap, 0, sec, stream));
  const long long threads = (long long)n * 64;
  split_mark_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(g, n, points, bitmap, new_points);
  NSVF_LAUNCH_OK("split_mark_kernel");
  scan_block_sums_kernel<<<n_blocks, kScanThreads, 0, stream>>>(n_words, bitmap, block_sums);
  NSVF_LAUNCH_OK("scan_block_sums_kernel");
  scan_offsets_kernel<<<1, kScanThreads, 0, stream>>>(n_blocks, block_sums, n_keys);
  NSVF_LAUNCH_OK("scan_offsets_kernel");
  scan_words_kernel<<<n_blocks, kScanThreads, 0, stream>>>(n_words, bitmap, block_sums, word_rank);
  NSVF_LAUNCH_OK("scan_words_kernel");
  return 0;
}

// Phase 2 (same workspace, untouched since phase 1): new_feats i32 [8n,8], parent i32 [Kc'] (scratch, any
// content), new_keys i32 [Kc',3] (optional), new_values f32 [Kc',D] (optional together with feats/values).
extern "C" int nsvf_split_emit(nsvf_stream_t stream_, int n, int D, const float* points, const int* feats,
                               const float* values, float half_voxel, const float* pmin, const int* max_coord,
                               int n_keys, int* new_feats, int* parent, int* new_keys, float* new_values,
                               void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SplitGrid g;
  long long n_words;
  if (split_grid(pmin, max_coord, half_voxel, g, n_words)) return 1;
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= nsvf_split_workspace_bytes(max_coord),
               "split: workspace too small");
  const long long sec = (n_words * 4 + 127) / 128 * 128;
  const unsigned* bitmap = (const unsigned*)workspace;
  const int* word_rank = (const int*)((const char*)workspace + sec);
  NSVF_CUDA_OK(cudaMemsetAsync(parent, 0x7f, sizeof(int) * (size_t)n_keys, stream));   // 0x7f7f7f7f > any voxel index
  const long long threads = (long long)n * 64;
  split_feats_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(g, n, points, bitmap, word_rank, new_feats,
                                                                           parent);
  NSVF_LAUNCH_OK("split_feats_kernel");
  if (new_values != nullptr || new_keys != nullptr) {
    NSVF_REQUIRE(new_values == nullptr || (feats != nullptr && values != nullptr && D > 0), "split: values need feats");
    split_values_kernel<<<(unsigned)(((long long)n * 32 + 255) / 256), 256, 0, stream>>>(
        g, n, new_values ? D : 0, points, feats, values, bitmap, word_rank, parent, new_keys, new_values);
    NSVF_LAUNCH_OK("split_values_kernel");
  }
  return 0;
}

extern "C" int nsvf_prune_lattice_embed(nsvf_stream_t stream_, int nv, int v0, int bits, int D, const int* feats,
                                        const float* centres, const float* values, float voxel_size, float* out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(nv >= 0 && bits >= 2 && bits <= 64, "prune_lattice_embed: bad sizes");
  NSVF_REQUIRE(D == 32, "prune_lattice_embed: only voxel_embed_dim == 32 is implemented (got %d)", D);
  if (nv == 0) return 0;
  const int L = bits * bits * bits;
  int gx = (4 * num_sms() + nv - 1) / nv;   // ~4 CTAs per SM in total; few CTAs per voxel amortise the row loads
  if (gx > (L + 31) / 32) gx = (L + 31) / 32;
  if (gx < 1) gx = 1;
  dim3 grid(gx, nv);
  prune_lattice_embed_d32_kernel<<<grid, 256, 0, stream>>>(nv, v0, bits, feats, centres, values, voxel_size,
                                                           voxel_size * 0.5f, out);
  NSVF_LAUNCH_OK("prune_lattice_embed_d32_kernel");
  return 0;
}

extern "C" int nsvf_prune_keep(nsvf_stream_t stream_, int nv, int L, const float* sigma, float th,
                               unsigned char* keep, float* min_score) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(nv >= 0 && L > 0, "prune_keep: bad sizes");
  if (nv == 0) return 0;
  prune_keep_kernel<<<(unsigned)(((long long)nv * 32 + 255) / 256), 256, 0, stream>>>(nv, L, sigma, th, keep, min_score);
  NSVF_LAUNCH_OK("prune_keep_kernel");
  return 0;
}
