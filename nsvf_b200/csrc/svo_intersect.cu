// Ray / sparse-voxel-octree intersection for sm_100a.
//
// Replaces fairnr/clib/src/intersect_gpu.cu:170-237 (svo_intersect_point_kernel) and its host glue
// fairnr/clib/src/intersect.cpp:84-112.  The reference emits leaves in its DFS order (children pushed
// 0..7, popped LIFO, root = node T-1) and truncates at n_max, so the traversal ORDER is part of the
// integer contract; we keep exactly that order and the bit-identical slab test (common.cuh) and change
// everything around it:
//   * a prep kernel packs each node into one 64-byte record {c-hv, c+hv, leaf flag, 8 children}
//     (reference: 3 + 9 scattered 4-byte loads per visit, hv recomputed per visit);
//   * 1/dir is computed once per ray (reference: 3 MUFU per node visit);
//   * the DFS stack is an uninitialised per-thread array (reference zero-fills 1 KiB per ray);
//   * output rows are pre-filled (-1 / 0) with coalesced stores by the whole CTA, then hits are
//     written in place (reference: 3 cudaMemsets + an uncoalesced per-thread -1 loop);
//   * TIGHT internal boxes.  The reference octree is loose (node half-extent 2^(depth+1) against a child offset of
//     2^(depth-1): every node box is twice its true size per axis), so most visited nodes are false positives.  A leaf
//     is emitted iff it and all its ancestors are hit; when every node box encloses its children's boxes (checked on
//     the device for each call) the ancestor tests are implied by the leaf test (monotone rounding, common.cuh), so
//     replacing the internal boxes by the exact union of their leaves' boxes changes neither the emitted set nor
//     the DFS order — it only prunes subtrees without hit leaves.  Unions are built per call by one climb per leaf
//     with float atomic min/max.  Irregular rays (NaN paths) and trees that fail the check use the loose boxes.
#include <cstdlib>

#include "common.cuh"
#include "nsvf_b200.h"
#include "voxel_grid.cuh"

namespace nsvf {

constexpr int kSvoStack = 256;   // reference bound (intersect_gpu.cu:205,209)
constexpr int kSvoThreads = 128;

struct __align__(16) SvoNode {
  float lo[3];
  float hi[3];
  int leaf;   // children[k][8] == 1
  int pad;
  int child[8];
};
static_assert(sizeof(SvoNode) == 64, "SvoNode must be one 64-byte record");

__global__ void svo_pack_kernel(const float* __restrict__ points, const int* __restrict__ children,
                                long long tree_stride_nodes, int T, float half_voxel,
                                SvoNode* __restrict__ nodes) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const float* p = points + ((long long)blockIdx.y * tree_stride_nodes + k) * 3;
  const int* ch = children + ((long long)blockIdx.y * tree_stride_nodes + k) * 9;
  SvoNode nd;
  int size = ch[8];
  // reference: half_voxel * float(children[k * 9 + 8])  (intersect_gpu.cu:217)
  float hv = __fmul_rn(half_voxel, (float)size);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float c = p[a];
    nd.lo[a] = __fsub_rn(c, hv);
    nd.hi[a] = __fadd_rn(c, hv);
  }
  nd.leaf = (size == 1);
  nd.pad = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) nd.child[u] = ch[u];
  int4* dst = reinterpret_cast<int4*>(nodes + (long long)blockIdx.y * T + k);
  const int4* src = reinterpret_cast<const int4*>(&nd);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = src[q];
}

// parent pointers + enclosure check of the reference's (loose) boxes; tight boxes start empty for internal nodes
__global__ void svo_prepare_kernel(const SvoNode* __restrict__ loose, int T, SvoNode* __restrict__ tight,
                                   int* __restrict__ parent, int* __restrict__ flag_bad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const SvoNode* ln = loose + (long long)blockIdx.y * T;
  SvoNode* tn = tight + (long long)blockIdx.y * T;
  int* par = parent + (long long)blockIdx.y * T;
  SvoNode nd = ln[k];
  if (!nd.leaf) {
    bool bad = false;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = nd.child[u];
      if (c > -1) {
        if (c >= T || c == k) { bad = true; continue; }
        const int old = atomicCAS(par + c, -1, k);
        if (old != -1 && old != k) bad = true;   // reachable from two parents: not a tree, keep the loose boxes
        const SvoNode cn = ln[c];
#pragma unroll
        for (int a = 0; a < 3; ++a) bad |= !(nd.lo[a] <= cn.lo[a]) || !(nd.hi[a] >= cn.hi[a]);
      }
    }
    if (bad) atomicExch(flag_bad + blockIdx.y, 1);
#pragma unroll
    for (int a = 0; a < 3; ++a) { nd.lo[a] = INFINITY; nd.hi[a] = -INFINITY; }
  }
  tn[k] = nd;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// every leaf climbs to the root and folds its exact box into its ancestors' unions
__global__ void svo_climb_kernel(int T, SvoNode* __restrict__ tight, const int* __restrict__ parent,
                                 int* __restrict__ flag_bad, int* __restrict__ leaf_count /* optional, zeroed */) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  SvoNode* tn = tight + (long long)blockIdx.y * T;
  const int* par = parent + (long long)blockIdx.y * T;
  if (!tn[k].leaf) return;
  float lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { lo[a] = tn[k].lo[a] + 0.0f; hi[a] = tn[k].hi[a] + 0.0f; }   // +0: canonical zero
  int* cnt = leaf_count != nullptr ? leaf_count + (long long)blockIdx.y * T : nullptr;
  if (cnt != nullptr) cnt[k] = 1;
  int p = par[k], steps = 0;
  while (p >= 0) {
    if (++steps > 64) { atomicExch(flag_bad + blockIdx.y, 1); break; }   // cycle or absurd depth
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomic_min_float(&tn[p].lo[a], lo[a]);
      atomic_max_float(&tn[p].hi[a], hi[a]);
    }
    if (cnt != nullptr) atomicAdd(cnt + p, 1);     // leaves below p
    p = par[p];
  }
}

// DFS emission rank of every leaf that is reachable from the root (node T-1), -1 for all other nodes.  The reference
// pushes the children 0..7 of a hit node and pops the last one first (intersect_gpu.cu:219-228), so among siblings the
// higher slot is visited first and a whole subtree is finished before the next sibling starts: the rank of a leaf is
// the number of leaves in the subtrees of its ancestors' higher-slot siblings, summed along its path to the root.
__global__ void svo_rank_kernel(int T, const SvoNode* __restrict__ loose, const int* __restrict__ parent,
                                const int* __restrict__ leaf_count, int* __restrict__ rank) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const SvoNode* ln = loose + (long long)blockIdx.y * T;
  const int* par = parent + (long long)blockIdx.y * T;
  const int* cnt = leaf_count + (long long)blockIdx.y * T;
  int r = -1;
  if (ln[k].leaf) {
    r = 0;
    int node = k, p = par[k], steps = 0;
    while (p >= 0 && ++steps <= 64) {
      bool seen = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = ln[p].child[u];
        if (c == node) { seen = true; continue; }
        if (seen && c > -1 && c < T) r += cnt[c];      // slots above the one we came from
      }
      node = p;
      p = par[p];
    }
    if (node != T - 1) r = -1;                        // not connected to the root: the traversal never gets there
  }
  rank[(long long)blockIdx.y * T + k] = r;
}

// widen the unions by one ulp (strict enclosure); internal nodes without any leaf can never be hit
__global__ void svo_finish_kernel(int T, SvoNode* __restrict__ tight) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  SvoNode* nd = tight + (long long)blockIdx.y * T + k;
  if (nd->leaf) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float l = nd->lo[a], h = nd->hi[a];
    if (l <= h) { nd->lo[a] = nextafterf(l, -INFINITY); nd->hi[a] = nextafterf(h, INFINITY); }
    else { nd->lo[a] = 3.0e38f; nd->hi[a] = 3.0e38f; }
  }
}

__global__ void __launch_bounds__(kSvoThreads)
svo_intersect_kernel(const SvoNode* __restrict__ nodes_all, const SvoNode* __restrict__ tight_all,
                     const int* __restrict__ flag_bad, int T, long long rays_per_tree, int n_max,
                     const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                     int* __restrict__ out_idx, float* __restrict__ out_min, float* __restrict__ out_max,
                     int* __restrict__ overflow_flag, const int* __restrict__ walk_active,
                     const unsigned char* __restrict__ defer) {
  // walk_active[tree] != 0: the lattice walk (voxel_grid.cu) has answered every ray of this tree except those it
  // marked in defer[]; only they are traversed here, and the rows of the others are left alone
  const bool only_deferred = walk_active != nullptr && walk_active[blockIdx.y] != 0;
  const SvoNode* loose = nodes_all + (long long)blockIdx.y * T;
  const bool use_tight = tight_all != nullptr && flag_bad[blockIdx.y] == 0;
  const SvoNode* tight = use_tight ? tight_all + (long long)blockIdx.y * T : loose;
  const long long ray_base = (long long)blockIdx.y * rays_per_tree;

  for (long long tile = (long long)blockIdx.x * kSvoThreads; tile < rays_per_tree;
       tile += (long long)gridDim.x * kSvoThreads) {
    const long long tile_rays = min((long long)kSvoThreads, rays_per_tree - tile);
    // 1) coalesced pre-fill of this tile's rows
    if (!only_deferred) {
      const long long base = (ray_base + tile) * n_max;
      const long long cells = tile_rays * n_max;
      for (long long c = threadIdx.x; c < cells; c += kSvoThreads) {
        out_idx[base + c] = -1;
        out_min[base + c] = 0.0f;
        out_max[base + c] = 0.0f;
      }
    }
    __syncthreads();
    // 2) one thread per ray: the reference DFS
    if (threadIdx.x < tile_rays && (!only_deferred || defer[ray_base + tile + threadIdx.x] != 0)) {
      const long long ray = ray_base + tile + threadIdx.x;
      if (only_deferred) {     // (rare) this ray's row was not pre-filled
        for (int c = 0; c < n_max; ++c) { out_idx[ray * n_max + c] = -1; out_min[ray * n_max + c] = 0.0f; out_max[ray * n_max + c] = 0.0f; }
      }
      const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
      const float ix = ref_rcp(ray_dir[ray * 3 + 0]), iy = ref_rcp(ray_dir[ray * 3 + 1]),
                  iz = ref_rcp(ray_dir[ray * 3 + 2]);
      const bool regular = regular_component(ox, ix) && regular_component(oy, iy) && regular_component(oz, iz);
      const SvoNode* nodes = regular ? tight : loose;   // NaN paths keep the reference's own boxes
      const long long row = ray * n_max;
      int stack[kSvoStack];
      int ptr = 0, cnt = 0;
      stack[0] = T - 1;  // ROOT node is always the last
      while (ptr > -1 && cnt < n_max) {
        const int k = stack[ptr--];
        const int4* rec = reinterpret_cast<const int4*>(nodes + k);
        const int4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
        const float lx = __int_as_float(q0.x), ly = __int_as_float(q0.y), lz = __int_as_float(q0.z);
        const float hx = __int_as_float(q0.w), hy = __int_as_float(q1.x), hz = __int_as_float(q1.y);
        float tn, tf;
        bool hit;
        if (regular) {
          // fmin/fmax ordering == the reference's swap when no NaN can occur
          float a0 = __fmul_rn(__fsub_rn(lx, ox), ix), b0 = __fmul_rn(__fsub_rn(hx, ox), ix);
          float a1 = __fmul_rn(__fsub_rn(ly, oy), iy), b1 = __fmul_rn(__fsub_rn(hy, oy), iy);
          float a2 = __fmul_rn(__fsub_rn(lz, oz), iz), b2 = __fmul_rn(__fsub_rn(hz, oz), iz);
          tn = fmaxf(fmaxf(fmaxf(0.0f, fminf(a0, b0)), fminf(a1, b1)), fminf(a2, b2));
          tf = fminf(fminf(fminf(100000.0f, fmaxf(a0, b0)), fmaxf(a1, b1)), fmaxf(a2, b2));
          hit = tn <= tf;
        } else {
          hit = slab_exact(ox, oy, oz, ix, iy, iz, lx, ly, lz, hx, hy, hz, tn, tf);
        }
        if (!hit) continue;
        if (q1.z) {  // terminal node
          out_idx[row + cnt] = k;
          out_min[row + cnt] = tn;
          out_max[row + cnt] = tf;
          ++cnt;
          continue;
        }
        const int4 c0 = __ldg(rec + 2), c1 = __ldg(rec + 3);
        if (ptr + 8 >= kSvoStack) {  // reference: device assert((ptr < 256)); we flag and stop this ray
          atomicExch(overflow_flag, 1);
          break;
        }
        if (c0.x > -1) stack[++ptr] = c0.x;
        if (c0.y > -1) stack[++ptr] = c0.y;
        if (c0.z > -1) stack[++ptr] = c0.z;
        if (c0.w > -1) stack[++ptr] = c0.w;
        if (c1.x > -1) stack[++ptr] = c1.x;
        if (c1.y > -1) stack[++ptr] = c1.y;
        if (c1.z > -1) stack[++ptr] = c1.z;
        if (c1.w > -1) stack[++ptr] = c1.w;
      }
    }
    __syncthreads();
  }
}

}  // namespace nsvf

using namespace nsvf;

extern "C" size_t nsvf_svo_workspace_bytes(int T, int n_trees) {
  if (T <= 0 || n_trees <= 0) return 0;
  // [flags 128 B + n_trees ints][loose nodes][tight nodes][parent]
  const size_t flags = 128 + ((size_t)n_trees * 4 + 127) / 128 * 128;
  return flags + (size_t)T * n_trees * (2 * sizeof(SvoNode) + sizeof(int)) + 256;
}

extern "C" int nsvf_svo_intersect(nsvf_stream_t stream_, int b, int T, int m, float voxelsize, int n_max,
                                  const float* ray_start, const float* ray_dir, const float* points,
                                  const int* children, long long tree_batch_stride_nodes, int* idx,
                                  float* min_depth, float* max_depth, void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSVF_REQUIRE(b >= 0 && T >= 0 && m >= 0 && n_max >= 0, "svo_intersect: negative size");
  if (b == 0 || m == 0 || n_max == 0) return 0;
  const long long rays = (long long)b * m;
  NSVF_REQUIRE(T > 0, "svo_intersect: empty octree (the reference reads node T-1 as the root)");
  NSVF_REQUIRE(tree_batch_stride_nodes == 0 || tree_batch_stride_nodes >= T,
               "svo_intersect: tree_batch_stride_nodes must be 0 (shared octree) or >= T");
  const int n_trees = tree_batch_stride_nodes == 0 ? 1 : b;
  const size_t need = nsvf_svo_workspace_bytes(T, n_trees);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= need, "svo_intersect: workspace too small (%zu < %zu bytes)",
               workspace_bytes, need);
  NSVF_REQUIRE(((uintptr_t)workspace & 127) == 0, "svo_intersect: workspace must be 128-byte aligned");
  const size_t flag_bytes = 128 + ((size_t)n_trees * 4 + 127) / 128 * 128;
  int* flag = (int*)workspace;                     // [0] = stack overflow, [32..] = per-tree "not enclosing"
  int* flag_bad = flag + 32;
  SvoNode* nodes = (SvoNode*)((char*)workspace + flag_bytes);
  SvoNode* tight = nodes + (size_t)T * n_trees;
  int* parent = (int*)(tight + (size_t)T * n_trees);
  NSVF_CUDA_OK(cudaMemsetAsync(flag, 0, flag_bytes, stream));
  NSVF_CUDA_OK(cudaMemsetAsync(parent, 0xff, sizeof(int) * (size_t)T * n_trees, stream));
  const float half_voxel = voxelsize * 0.5f;
  {
    dim3 grid((T + 255) / 256, n_trees);
    svo_pack_kernel<<<grid, 256, 0, stream>>>(points, children, tree_batch_stride_nodes, T, half_voxel, nodes);
    NSVF_LAUNCH_OK("svo_pack_kernel");
    svo_prepare_kernel<<<grid, 256, 0, stream>>>(nodes, T, tight, parent, flag_bad);
    NSVF_LAUNCH_OK("svo_prepare_kernel");
    svo_climb_kernel<<<grid, 256, 0, stream>>>(T, tight, parent, flag_bad, nullptr);
    NSVF_LAUNCH_OK("svo_climb_kernel");
    svo_finish_kernel<<<grid, 256, 0, stream>>>(T, tight);
    NSVF_LAUNCH_OK("svo_finish_kernel");
  }
  const long long rays_per_tree = n_trees == 1 ? rays : m;
  long long want = (rays_per_tree + kSvoThreads - 1) / kSvoThreads;
  long long cap = (long long)num_sms() * 8;
  if (n_trees > 1) cap = (cap + n_trees - 1) / n_trees;
  int gx = (int)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_trees);
  NSVF_TIMED_LAUNCH("svo_intersect_kernel", stream, (svo_intersect_kernel<<<grid, kSvoThreads, 0, stream>>>(
                        nodes, getenv("NSVF_SVO_LOOSE") ? nullptr : tight, flag_bad, T, rays_per_tree, n_max, ray_start,
                        ray_dir, idx, min_depth, max_depth, flag, nullptr, nullptr)));
  return 0;
}

// ---- octree intersection with the encoder's post-processing fused in ----------------------------------------------
// nsvf_svo_intersect + nsvf_sort_hits_by_depth in one call, and on a different algorithm whenever the tree allows it:
// once every node box is known to enclose its children's boxes, a leaf is reported iff its OWN slab test hits (file
// header), i.e. the answer is that of the voxel-set query over the leaves — and the leaves are a lattice, so the rays
// walk it (voxel_grid.cu: a few hundred cell lookups instead of a DFS over eight-way nodes, hits already in depth
// order).  What the DFS contributed besides the hit set is reproduced through the leaves' DFS ranks: ties in depth are
// ordered by rank and a row that overflows n_max keeps the n_max smallest ranks, exactly what "first n_max leaves in
// DFS order, then a stable sort by depth" gives.  Trees that fail the checks, leaves off a common lattice and rays with
// NaN paths go through the traversal kernel + sort as before; every decision is taken on the device.
// workspace of the tree: [nsvf_svo_workspace_bytes layout][leaf counts][DFS ranks][lattice]; per-call ray scratch:
// [walk_active, one int per tree, 128-byte padded][defer, one byte per ray]
static size_t svo_tree_offsets(int T, int n_trees, size_t* off_count, size_t* off_rank, size_t* off_grid) {
  size_t o = (nsvf_svo_workspace_bytes(T, n_trees) + 127) / 128 * 128;
  *off_count = o; o += ((size_t)T * n_trees * 4 + 127) / 128 * 128;
  *off_rank = o;  o += ((size_t)T * n_trees * 4 + 127) / 128 * 128;
  *off_grid = o;  o += voxel_grid_bytes(T) * (size_t)n_trees;
  return o;
}
static size_t svo_scratch_bytes(int n_trees, long long rays) {
  return ((size_t)n_trees * 4 + 127) / 128 * 128 + ((size_t)(rays > 0 ? rays : 0) + 127) / 128 * 128;
}

extern "C" size_t nsvf_svo_sorted_workspace_bytes(int T, int n_trees, long long rays) {
  if (T <= 0 || n_trees <= 0) return 0;
  size_t a, b, c;
  return svo_tree_offsets(T, n_trees, &a, &b, &c) + svo_scratch_bytes(n_trees, rays);
}

enum { kSvoBuild = 1, kSvoTraverse = 2 };

static int svo_sorted_run(cudaStream_t stream, int phase, int b, int T, int m, float voxelsize, int n_max,
                          float empty_depth, const float* ray_start, const float* ray_dir, const float* points,
                          const int* children, long long tree_batch_stride_nodes, int* idx, float* min_depth,
                          float* max_depth, unsigned char* hits, void* tree_ws, size_t tree_ws_bytes, void* scratch,
                          size_t scratch_bytes) {
  NSVF_REQUIRE(b >= 0 && T >= 0 && m >= 0 && n_max >= 0, "svo_intersect_sorted: negative size");
  const bool traverse = (phase & kSvoTraverse) != 0;
  if (traverse && (b == 0 || m == 0 || n_max == 0)) return 0;
  const long long rays = (long long)b * m;
  NSVF_REQUIRE(T > 0, "svo_intersect_sorted: empty octree (the reference reads node T-1 as the root)");
  NSVF_REQUIRE(tree_batch_stride_nodes == 0 || tree_batch_stride_nodes >= T,
               "svo_intersect_sorted: tree_batch_stride_nodes must be 0 (shared octree) or >= T");
  const int n_trees = tree_batch_stride_nodes == 0 ? 1 : b;
  size_t off_count, off_rank, off_grid;
  const size_t need = svo_tree_offsets(T, n_trees, &off_count, &off_rank, &off_grid);
  NSVF_REQUIRE(tree_ws != nullptr && tree_ws_bytes >= need, "svo_intersect_sorted: workspace too small (%zu < %zu bytes)",
               tree_ws_bytes, need);
  NSVF_REQUIRE(((uintptr_t)tree_ws & 127) == 0, "svo_intersect_sorted: workspace must be 128-byte aligned");
  char* ws = (char*)tree_ws;
  const size_t flag_bytes = 128 + ((size_t)n_trees * 4 + 127) / 128 * 128;
  int* flag = (int*)ws;
  int* flag_bad = flag + 32;
  SvoNode* nodes = (SvoNode*)(ws + flag_bytes);
  SvoNode* tight = nodes + (size_t)T * n_trees;
  int* parent = (int*)(tight + (size_t)T * n_trees);
  int* leaf_count = (int*)(ws + off_count);
  int* rank = (int*)(ws + off_rank);
  const size_t grid_set_bytes = voxel_grid_bytes(T);
  unsigned char* grid_ws = grid_set_bytes ? (unsigned char*)(ws + off_grid) : nullptr;
  const long long pts_stride = tree_batch_stride_nodes * 3;
  if (phase & kSvoBuild) {
    NSVF_CUDA_OK(cudaMemsetAsync(flag, 0, flag_bytes, stream));
    NSVF_CUDA_OK(cudaMemsetAsync(parent, 0xff, sizeof(int) * (size_t)T * n_trees, stream));
    NSVF_CUDA_OK(cudaMemsetAsync(leaf_count, 0, sizeof(int) * (size_t)T * n_trees, stream));
    const float half_voxel = voxelsize * 0.5f;
    dim3 grid((T + 255) / 256, n_trees);
    svo_pack_kernel<<<grid, 256, 0, stream>>>(points, children, tree_batch_stride_nodes, T, half_voxel, nodes);
    NSVF_LAUNCH_OK("svo_pack_kernel");
    svo_prepare_kernel<<<grid, 256, 0, stream>>>(nodes, T, tight, parent, flag_bad);
    NSVF_LAUNCH_OK("svo_prepare_kernel");
    svo_climb_kernel<<<grid, 256, 0, stream>>>(T, tight, parent, flag_bad, leaf_count);
    NSVF_LAUNCH_OK("svo_climb_kernel");
    svo_finish_kernel<<<grid, 256, 0, stream>>>(T, tight);
    NSVF_LAUNCH_OK("svo_finish_kernel");
    if (grid_ws != nullptr) {
      svo_rank_kernel<<<grid, 256, 0, stream>>>(T, nodes, parent, leaf_count, rank);
      NSVF_LAUNCH_OK("svo_rank_kernel");
      if (voxel_grid_build(stream, n_trees, T, points, pts_stride, voxelsize, grid_ws, grid_set_bytes, rank, T)) return 1;
    }
  }
  if (!traverse) return 0;
  NSVF_REQUIRE(scratch != nullptr && scratch_bytes >= svo_scratch_bytes(n_trees, rays) && ((uintptr_t)scratch & 127) == 0,
               "svo_intersect_sorted: ray scratch too small or not 128-byte aligned (%zu < %zu bytes)", scratch_bytes,
               svo_scratch_bytes(n_trees, rays));
  int* walk_active = (int*)scratch;
  unsigned char* defer = (unsigned char*)scratch + ((size_t)n_trees * 4 + 127) / 128 * 128;
  NSVF_CUDA_OK(cudaMemsetAsync(walk_active, 0, sizeof(int) * (size_t)n_trees, stream));
  const long long rays_per_tree = n_trees == 1 ? rays : m;
  if (grid_ws != nullptr) {
    WalkOctree oct;
    oct.rank = rank; oct.rank_stride = T; oct.veto = flag_bad; oct.active = walk_active; oct.defer = defer;
    if (voxel_grid_walk(stream, 1, grid_ws, grid_set_bytes, n_trees, T, points, pts_stride, voxelsize, rays_per_tree,
                        n_max, empty_depth, ray_start, ray_dir, idx, min_depth, max_depth, hits, &oct))
      return 1;
  }
  long long want = (rays_per_tree + kSvoThreads - 1) / kSvoThreads;
  long long cap = (long long)num_sms() * 8;
  if (n_trees > 1) cap = (cap + n_trees - 1) / n_trees;
  int gx = (int)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_trees);
  NSVF_TIMED_LAUNCH("svo_intersect_kernel", stream, (svo_intersect_kernel<<<grid, kSvoThreads, 0, stream>>>(
                        nodes, getenv("NSVF_SVO_LOOSE") ? nullptr : tight, flag_bad, T, rays_per_tree, n_max, ray_start,
                        ray_dir, idx, min_depth, max_depth, flag, grid_ws != nullptr ? walk_active : nullptr, defer)));
  return sort_hits_run(stream, rays, n_max, empty_depth, idx, min_depth, max_depth, hits,
                       grid_ws != nullptr ? walk_active : nullptr, rays_per_tree, defer);
}

extern "C" int nsvf_svo_intersect_sorted(nsvf_stream_t stream, int b, int T, int m, float voxelsize, int n_max,
                                         float empty_depth, const float* ray_start, const float* ray_dir,
                                         const float* points, const int* children, long long tree_batch_stride_nodes,
                                         int* idx, float* min_depth, float* max_depth, unsigned char* hits,
                                         void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(b >= 0 && T >= 0 && m >= 0 && n_max >= 0, "svo_intersect_sorted: negative size");
  if (b == 0 || m == 0 || n_max == 0) return 0;
  NSVF_REQUIRE(T > 0, "svo_intersect_sorted: empty octree (the reference reads node T-1 as the root)");
  const int n_trees = tree_batch_stride_nodes == 0 ? 1 : b;
  size_t a, c, d;
  const size_t tree_bytes = svo_tree_offsets(T, n_trees, &a, &c, &d);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= tree_bytes + svo_scratch_bytes(n_trees, (long long)b * m),
               "svo_intersect_sorted: workspace too small (%zu < %zu bytes)", workspace_bytes,
               tree_bytes + svo_scratch_bytes(n_trees, (long long)b * m));
  return svo_sorted_run((cudaStream_t)stream, kSvoBuild | kSvoTraverse, b, T, m, voxelsize, n_max, empty_depth, ray_start,
                        ray_dir, points, children, tree_batch_stride_nodes, idx, min_depth, max_depth, hits, workspace,
                        tree_bytes, (char*)workspace + tree_bytes, workspace_bytes - tree_bytes);
}

extern "C" int nsvf_svo_prepare(nsvf_stream_t stream, int n_trees, int T, float voxelsize, const float* points,
                                const int* children, long long tree_batch_stride_nodes, void* workspace,
                                size_t workspace_bytes) {
  NSVF_REQUIRE(n_trees >= 1 && (tree_batch_stride_nodes != 0 || n_trees == 1),
               "svo_prepare: a shared octree (tree_batch_stride_nodes 0) is one tree");
  return svo_sorted_run((cudaStream_t)stream, kSvoBuild, n_trees, T, 0, voxelsize, 0, 0.0f, nullptr, nullptr, points, children,
                        tree_batch_stride_nodes, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, nullptr, 0);
}

extern "C" size_t nsvf_svo_ray_scratch_bytes(int n_trees, long long rays) {
  return n_trees <= 0 ? 0 : svo_scratch_bytes(n_trees, rays);
}

extern "C" int nsvf_svo_intersect_sorted_prepared(nsvf_stream_t stream, int b, int T, int m, float voxelsize, int n_max,
                                                  float empty_depth, const float* ray_start, const float* ray_dir,
                                                  const float* points, const int* children,
                                                  long long tree_batch_stride_nodes, int* idx, float* min_depth,
                                                  float* max_depth, unsigned char* hits, const void* workspace,
                                                  size_t workspace_bytes, void* ray_scratch, size_t ray_scratch_bytes) {
  return svo_sorted_run((cudaStream_t)stream, kSvoTraverse, b, T, m, voxelsize, n_max, empty_depth, ray_start, ray_dir,
                        points, children, tree_batch_stride_nodes, idx, min_depth, max_depth, hits,
                        const_cast<void*>(workspace), workspace_bytes, ray_scratch, ray_scratch_bytes);
}
