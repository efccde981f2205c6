// Ray / voxel-AABB intersection for sm_100a: first n_max hit voxels per ray in ascending voxel index.
//
// Replaces fairnr/clib/src/intersect_gpu.cu:125-167 (aabb_intersect_point_kernel, one thread per ray
// scanning all n voxels) and its host glue fairnr/clib/src/intersect.cpp:49-75.
//
// Design (not a port):
//   * The reference result is "the n_max smallest voxel indices whose slab test hits".  We keep the
//     slab test bit-identical (common.cuh) but never scan all voxels: an implicit 32-ary hierarchy of
//     enclosing boxes is built over the voxels IN INDEX ORDER (node j of level l covers voxels
//     [j*32^l, (j+1)*32^l)), so a depth-first walk in child order emits hits in ascending voxel index
//     and can stop at n_max exactly like the reference loop does.
//   * One warp owns one ray.  Each step tests the 32 children of a node, one per lane, with coalesced
//     SoA loads; __ballot_sync gives the hit mask, __popc of the lower lanes gives each hit's output
//     rank (ballot/prefix compaction).  Control flow is warp-uniform: no divergence.
//   * The upper levels of the hierarchy (everything that fits NSVF_AABB_SMEM_NODES) are staged into
//     shared memory once per persistent CTA with TMA bulk copies (cp.async.bulk + mbarrier).
//   * Hits are collected in shared memory and each ray's row [n_max] x {idx,min,max} is written once,
//     coalesced, including the -1 / 0 fill that the reference gets from torch::zeros + a per-thread loop.
#include "common.cuh"
#include "nsvf_b200.h"

namespace nsvf {

constexpr int kAabbMaxLevels = 7;          // 32^7 > 2^31 voxels
constexpr int kAabbSmemNodes = 4096;       // nodes staged per CTA: 6 * 4096 * 4 B = 96 KiB
constexpr int kAabbWarps = 8;

struct AabbTree {
  const float* box;    // SoA: 6 arrays of `total` floats: lo.x lo.y lo.z hi.x hi.y hi.z
  int total;           // padded node count over all levels (multiple of 32)
  int nlevels;         // level 0 = voxels
  int cnt[kAabbMaxLevels];
  int off[kAabbMaxLevels];
  int stage_from;      // first array index staged in shared memory (== total: nothing staged)
};

struct AabbLayout {
  int nlevels, total, stage_from;
  int cnt[kAabbMaxLevels], off[kAabbMaxLevels];
};

static AabbLayout aabb_layout(int n) {
  AabbLayout L{};
  int c = n, o = 0, l = 0;
  for (;;) {
    L.cnt[l] = c;
    L.off[l] = o;
    o += (c + 31) / 32 * 32;
    ++l;
    if (c <= 32) break;
    c = (c + 31) / 32;
  }
  L.nlevels = l;
  L.total = o;
  // stage the largest suffix of levels that fits
  L.stage_from = L.total;
  for (int k = l - 1; k >= 0; --k) {
    if (L.total - L.off[k] <= kAabbSmemNodes) L.stage_from = L.off[k];
    else break;
  }
  return L;
}

// ---- hierarchy build ----------------------------------------------------------------------------
// level 0 + level 1 in one pass: one warp per level-1 node reads 32 voxel centres, writes their exact
// boxes (c - hv, c + hv: the reference's first rounding) and the strictly enclosing parent box.
__global__ void aabb_build_l01_kernel(const float* __restrict__ points, long long tree_stride_pts, int n,
                                      float half_voxel, float* __restrict__ box_all, long long tree_stride_box,
                                      int total, int off1, int has_l1) {
  const float* pts = points + (long long)blockIdx.y * tree_stride_pts;
  float* box = box_all + (long long)blockIdx.y * tree_stride_box;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int n1 = (n + 31) / 32;
  if (warp >= n1) return;
  int i = warp * 32 + lane;
  float lo[3], hi[3];
  bool valid = i < n;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float c = valid ? pts[(long long)i * 3 + a] : 0.0f;
    lo[a] = __fsub_rn(c, half_voxel);
    hi[a] = __fadd_rn(c, half_voxel);
    box[(long long)a * total + i] = valid ? lo[a] : 0.0f;
    box[(long long)(3 + a) * total + i] = valid ? hi[a] : 0.0f;
    if (!valid) { lo[a] = INFINITY; hi[a] = -INFINITY; }
  }
  if (!has_l1) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(NSVF_FULL_MASK, lo[a], s));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(NSVF_FULL_MASK, hi[a], s));
    }
  }
  if (lane < 3) {
    box[(long long)lane * total + off1 + warp] = nextafterf(lo[lane], -INFINITY);
    box[(long long)(3 + lane) * total + off1 + warp] = nextafterf(hi[lane], INFINITY);
  }
  if (warp == n1 - 1) {  // zero the level-1 padding (it is staged by TMA, never tested)
    int pad_end = (n1 + 31) / 32 * 32;
    for (int j = n1 + lane; j < pad_end; j += 32)
      for (int a = 0; a < 6; ++a) box[(long long)a * total + off1 + j] = 0.0f;
  }
}

// level l >= 2 from level l-1 (already strict): one warp per node.
__global__ void aabb_build_up_kernel(float* __restrict__ box_all, long long tree_stride_box, int total,
                                     int off_child, int cnt_child, int off_parent, int cnt_parent) {
  float* box = box_all + (long long)blockIdx.y * tree_stride_box;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= cnt_parent) return;
  int i = warp * 32 + lane;
  bool valid = i < cnt_child;
  float lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = valid ? box[(long long)a * total + off_child + i] : INFINITY;
    hi[a] = valid ? box[(long long)(3 + a) * total + off_child + i] : -INFINITY;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(NSVF_FULL_MASK, lo[a], s));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(NSVF_FULL_MASK, hi[a], s));
    }
  }
  if (lane < 3) {
    box[(long long)lane * total + off_parent + warp] = lo[lane];
    box[(long long)(3 + lane) * total + off_parent + warp] = hi[lane];
  }
  // zero the padding of the parent level so staged copies never carry uninitialised words
  if (warp == cnt_parent - 1) {
    int pad_end = (cnt_parent + 31) / 32 * 32;
    for (int j = cnt_parent + lane; j < pad_end; j += 32)
      for (int a = 0; a < 6; ++a) box[(long long)a * total + off_parent + j] = 0.0f;
  }
}

// ---- traversal ------------------------------------------------------------------------------------
struct AabbRay {
  float ox, oy, oz, ix, iy, iz;
  bool regular;
  int nx, ny, nz;  // near-array selectors for the sorted test: 0 -> lo is near, 3 -> hi is near
};

struct AabbWarp {
  const float* gbox;   // global SoA of this tree
  const float* sbox;   // shared SoA (array stride = sm_nodes)
  int sm_nodes;
  int* h_idx;          // per-warp hit buffers in shared memory
  float* h_min;
  float* h_max;
  int n_max;
};

// `tree` is the __grid_constant__ kernel parameter: cnt[L] / off[L] with a compile-time L are direct
// constant-bank operands, no registers.
template <int L>
__device__ __forceinline__ void aabb_descend(const AabbTree& tree, const AabbWarp& c, const AabbRay& r, int base,
                                             int& cnt) {
  const int lane = threadIdx.x & 31;
  const int i = base + lane;
  bool hit = false;
  float tn = 0.f, tf = 0.f;
  if (i < tree.cnt[L]) {
    const int pos = tree.off[L] + i;
    const bool staged = tree.off[L] >= tree.stage_from;  // uniform per level
    const float* bp = staged ? c.sbox + (pos - tree.stage_from) : c.gbox + pos;
    const int stride = staged ? c.sm_nodes : tree.total;
    if (r.regular) {
      const float nx = bp[(long long)r.nx * stride], fx = bp[(long long)(3 - r.nx) * stride];
      const float ny = bp[(long long)(1 + r.ny) * stride], fy = bp[(long long)(4 - r.ny) * stride];
      const float nz = bp[(long long)(2 + r.nz) * stride], fz = bp[(long long)(5 - r.nz) * stride];
      hit = slab_sorted(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, nx, ny, nz, fx, fy, fz, tn, tf);
    } else {
      const float lx = bp[0], ly = bp[(long long)stride], lz = bp[(long long)2 * stride];
      const float hx = bp[(long long)3 * stride], hy = bp[(long long)4 * stride], hz = bp[(long long)5 * stride];
      if (L == 0) hit = slab_exact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz, tn, tf);
      else hit = slab_enclosing(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz);
    }
  }
  unsigned m = __ballot_sync(NSVF_FULL_MASK, hit);
  if constexpr (L == 0) {
    if (hit) {
      const int rank = cnt + __popc(m & ((1u << lane) - 1u));
      if (rank < c.n_max) {
        c.h_idx[rank] = i;
        c.h_min[rank] = tn;
        c.h_max[rank] = tf;
      }
    }
    cnt += __popc(m);
  } else {
    while (m != 0u && cnt < c.n_max) {
      const int b = __ffs(m) - 1;
      m &= m - 1u;
      aabb_descend<L - 1>(tree, c, r, (base + b) * 32, cnt);
    }
  }
}

__global__ void __launch_bounds__(kAabbWarps * 32)
aabb_intersect_kernel(const __grid_constant__ AabbTree tree, long long tree_stride_box, long long rays_per_tree,
                      int n_max, const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                      int* __restrict__ out_idx, float* __restrict__ out_min, float* __restrict__ out_max) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int sm_nodes = tree.total - tree.stage_from;   // multiple of 32
  float* sbox = reinterpret_cast<float*>(smem_raw);
  int* hbuf = reinterpret_cast<int*>(smem_raw + (size_t)6 * sm_nodes * sizeof(float));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* gbox = tree.box + (long long)blockIdx.y * tree_stride_box;

  // stage the upper levels with TMA bulk copies: 6 arrays, one mbarrier
  if (sm_nodes > 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned bytes = (unsigned)sm_nodes * 4u;
      mbar_expect_tx(&bar, bytes * 6u);
#pragma unroll
      for (int a = 0; a < 6; ++a)
        tma_bulk_g2s(sbox + (size_t)a * sm_nodes, gbox + (long long)a * tree.total + tree.stage_from, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  }

  AabbWarp c;
  c.gbox = gbox;
  c.sbox = sbox;
  c.sm_nodes = sm_nodes;
  c.n_max = n_max;
  c.h_idx = hbuf + warp * 3 * n_max;
  c.h_min = reinterpret_cast<float*>(c.h_idx + n_max);
  c.h_max = c.h_min + n_max;

  const long long ray_base = (long long)blockIdx.y * rays_per_tree;
  for (long long rr = (long long)blockIdx.x * kAabbWarps + warp; rr < rays_per_tree;
       rr += (long long)gridDim.x * kAabbWarps) {
    const long long ray = ray_base + rr;
    // lanes 0..2 load the origin, 3..5 the direction; broadcast
    float v = 0.f;
    if (lane < 3) v = ray_start[ray * 3 + lane];
    else if (lane < 6) v = ray_dir[ray * 3 + (lane - 3)];
    AabbRay r;
    r.ox = __shfl_sync(NSVF_FULL_MASK, v, 0);
    r.oy = __shfl_sync(NSVF_FULL_MASK, v, 1);
    r.oz = __shfl_sync(NSVF_FULL_MASK, v, 2);
    r.ix = ref_rcp(__shfl_sync(NSVF_FULL_MASK, v, 3));
    r.iy = ref_rcp(__shfl_sync(NSVF_FULL_MASK, v, 4));
    r.iz = ref_rcp(__shfl_sync(NSVF_FULL_MASK, v, 5));
    r.regular = regular_component(r.ox, r.ix) && regular_component(r.oy, r.iy) && regular_component(r.oz, r.iz);
    r.nx = r.ix < 0.f ? 3 : 0;
    r.ny = r.iy < 0.f ? 3 : 0;
    r.nz = r.iz < 0.f ? 3 : 0;

    int cnt = 0;
    switch (tree.nlevels) {
      case 1: aabb_descend<0>(tree, c, r, 0, cnt); break;
      case 2: aabb_descend<1>(tree, c, r, 0, cnt); break;
      case 3: aabb_descend<2>(tree, c, r, 0, cnt); break;
      case 4: aabb_descend<3>(tree, c, r, 0, cnt); break;
      case 5: aabb_descend<4>(tree, c, r, 0, cnt); break;
      case 6: aabb_descend<5>(tree, c, r, 0, cnt); break;
      default: aabb_descend<6>(tree, c, r, 0, cnt); break;
    }
    cnt = min(cnt, n_max);
    __syncwarp();
    const long long row = ray * n_max;
    for (int l = lane; l < n_max; l += 32) {
      const bool ok = l < cnt;
      out_idx[row + l] = ok ? c.h_idx[l] : -1;
      out_min[row + l] = ok ? c.h_min[l] : 0.0f;
      out_max[row + l] = ok ? c.h_max[l] : 0.0f;
    }
    __syncwarp();
  }
}

static size_t aabb_tree_floats(int n) { return (size_t)6 * aabb_layout(n).total; }

}  // namespace nsvf

using namespace nsvf;

extern "C" size_t nsvf_aabb_workspace_bytes(int n, int n_trees) {
  if (n <= 0 || n_trees <= 0) return 0;
  return aabb_tree_floats(n) * sizeof(float) * (size_t)n_trees;
}

extern "C" int nsvf_aabb_intersect(nsvf_stream_t stream_, int b, int n, int m, float voxelsize, int n_max,
                                   const float* ray_start, const float* ray_dir, const float* points,
                                   long long points_batch_stride, int* idx, float* min_depth, float* max_depth,
                                   void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && n >= 0 && m >= 0 && n_max >= 0, "aabb_intersect: negative size");
  if (b == 0 || m == 0 || n_max == 0) return 0;
  const long long rays = (long long)b * m;
  if (n == 0) {  // nothing to hit: rows are all -1 / 0
    NSVF_CUDA_OK(cudaMemsetAsync(idx, 0xff, sizeof(int) * rays * n_max, stream));
    NSVF_CUDA_OK(cudaMemsetAsync(min_depth, 0, sizeof(float) * rays * n_max, stream));
    NSVF_CUDA_OK(cudaMemsetAsync(max_depth, 0, sizeof(float) * rays * n_max, stream));
    return 0;
  }
  NSVF_REQUIRE(points_batch_stride == 0 || points_batch_stride >= (long long)n * 3,
               "aabb_intersect: points_batch_stride must be 0 (shared voxel set) or >= 3*n");
  const int n_trees = points_batch_stride == 0 ? 1 : b;
  const size_t need = nsvf_aabb_workspace_bytes(n, n_trees);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= need,
               "aabb_intersect: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  NSVF_REQUIRE(((uintptr_t)workspace & 127) == 0, "aabb_intersect: workspace must be 128-byte aligned");

  AabbLayout L = aabb_layout(n);
  float* box = (float*)workspace;
  const long long tree_stride_box = (long long)6 * L.total;
  const float half_voxel = voxelsize * 0.5f;  // reference: float half_voxel = voxelsize * 0.5 (exact)

  {  // build
    int n1 = (n + 31) / 32;
    dim3 grid((n1 + 7) / 8, n_trees);
    aabb_build_l01_kernel<<<grid, 256, 0, stream>>>(points, points_batch_stride, n, half_voxel, box,
                                                    tree_stride_box, L.total, L.nlevels > 1 ? L.off[1] : 0,
                                                    L.nlevels > 1);
    NSVF_LAUNCH_OK("aabb_build_l01_kernel");
    for (int l = 2; l < L.nlevels; ++l) {
      dim3 g((L.cnt[l] + 7) / 8, n_trees);
      aabb_build_up_kernel<<<g, 256, 0, stream>>>(box, tree_stride_box, L.total, L.off[l - 1], L.cnt[l - 1],
                                                  L.off[l], L.cnt[l]);
      NSVF_LAUNCH_OK("aabb_build_up_kernel");
    }
  }

  AabbTree tree;
  tree.box = box;
  tree.total = L.total;
  tree.nlevels = L.nlevels;
  for (int l = 0; l < kAabbMaxLevels; ++l) { tree.cnt[l] = L.cnt[l]; tree.off[l] = L.off[l]; }
  tree.stage_from = L.stage_from;

  const long long rays_per_tree = n_trees == 1 ? rays : m;
  const int sm_nodes = L.total - L.stage_from;
  size_t smem = (size_t)6 * sm_nodes * sizeof(float) + (size_t)kAabbWarps * 3 * n_max * sizeof(int);
  NSVF_REQUIRE(smem <= 200 * 1024, "aabb_intersect: n_max=%d needs %zu B of shared memory", n_max, smem);
  static bool attr_set = false;
  if (!attr_set) {
    NSVF_CUDA_OK(cudaFuncSetAttribute(aabb_intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
    attr_set = true;
  }
  int blocks_per_sm = (int)((220 * 1024) / (smem + 1024));
  blocks_per_sm = blocks_per_sm < 1 ? 1 : (blocks_per_sm > 8 ? 8 : blocks_per_sm);
  long long want = (rays_per_tree + kAabbWarps - 1) / kAabbWarps;
  long long cap = (long long)num_sms() * blocks_per_sm;
  if (n_trees > 1) cap = (cap + n_trees - 1) / n_trees;
  int gx = (int)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_trees);
  aabb_intersect_kernel<<<grid, kAabbWarps * 32, smem, stream>>>(tree, tree_stride_box, rays_per_tree, n_max,
                                                                 ray_start, ray_dir, idx, min_depth, max_depth);
  NSVF_LAUNCH_OK("aabb_intersect_kernel");
  return 0;
}
