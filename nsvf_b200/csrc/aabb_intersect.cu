// Ray / voxel-AABB intersection for sm_100a: first n_max hit voxels per ray in ascending voxel index.
//
// Replaces fairnr/clib/src/intersect_gpu.cu:125-167 (aabb_intersect_point_kernel, one thread per ray
// scanning all n voxels) and its host glue fairnr/clib/src/intersect.cpp:49-75.
//
// Design (not a port):
//   * The reference result is "the n_max smallest voxel indices whose slab test hits".  We keep the slab test
//     bit-identical (common.cuh) but never scan all voxels: an implicit 8-ary hierarchy of enclosing boxes is built
//     over the voxels IN INDEX ORDER (node j of level l covers voxels [j*8^l, (j+1)*8^l); after a split the 8
//     children of a voxel are consecutive, so level-1 nodes are exactly the parent voxels).
//   * One warp owns one ray and walks the hierarchy breadth-first: the hit nodes of a level live in an ascending
//     list in shared memory; each step takes FOUR of them and tests their 4 x 8 children, one per lane (dense lane
//     use at every level), `__ballot_sync` + `__popc` append the hit children in ascending order to the next list
//     (ballot/prefix compaction).  At the leaf level the same step emits hits in ascending voxel index and stops at
//     n_max exactly like the reference's linear scan.  Control flow is warp-uniform: no divergence.
//   * The upper levels of the hierarchy (everything that fits kAabbSmemNodes) are staged into shared memory once
//     per persistent CTA with TMA bulk copies (cp.async.bulk + mbarrier).
//   * Hits are collected in shared memory and each ray's row [n_max] x {idx,min,max} is written once, coalesced,
//     including the -1 / 0 fill that the reference gets from torch::zeros + a per-thread loop; optionally sorted by
//     entry depth first (a register bitonic network), which replaces encoder.py:519-524.
//   * Few hundred voxels: a thread-per-ray scan over shared-memory boxes (aabb_small_kernel) beats any hierarchy.
#include <cstdlib>

#include "common.cuh"
#include "nsvf_b200.h"
#include "voxel_grid.cuh"

namespace nsvf {

constexpr int kAabbMaxLevels = 11;         // 8^10 * 32 > 2^31 voxels
constexpr int kAabbSmemNodes = 512;        // nodes staged per CTA (6 * 512 * 4 B = 12 KiB): the top levels only.  Measured
                                           // (C3 / C4 sorted intersection): 4096 nodes 4.46 / 8.86 ms, 512 nodes 4.25 / 7.57 ms —
                                           // the lower levels are served by L1 and the smaller footprint buys occupancy
constexpr int kAabbWarps = 8;
constexpr int kAabbListCap = 256;          // hit nodes kept per level and ray (continuation pass beyond that)

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
    c = (c + 7) / 8;
  }
  L.nlevels = l;
  L.total = o;
  L.stage_from = L.total;   // stage the largest suffix of levels that fits
  static const int smem_nodes = getenv("NSVF_AABB_SMEM_NODES") ? atoi(getenv("NSVF_AABB_SMEM_NODES")) : kAabbSmemNodes;
  for (int k = l - 1; k >= 0; --k) {
    if (L.total - L.off[k] <= smem_nodes) L.stage_from = L.off[k];
    else break;
  }
  return L;
}

// ---- hierarchy build ----------------------------------------------------------------------------
// level 0: exact voxel boxes (c - hv, c + hv: the reference's first rounding), padding zeroed
__global__ void aabb_build_leaves_kernel(const float* __restrict__ points, long long tree_stride_pts, int n,
                                         float half_voxel, float* __restrict__ box_all, long long tree_stride_box,
                                         int total) {
  const float* pts = points + (long long)blockIdx.y * tree_stride_pts;
  float* box = box_all + (long long)blockIdx.y * tree_stride_box;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pad = (n + 31) / 32 * 32;
  if (i >= n_pad) return;
  const bool valid = i < n;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float c = valid ? pts[(long long)i * 3 + a] : 0.0f;
    box[(long long)a * total + i] = valid ? __fsub_rn(c, half_voxel) : 0.0f;
    box[(long long)(3 + a) * total + i] = valid ? __fadd_rn(c, half_voxel) : 0.0f;
  }
}

// level l >= 1 from level l-1: one thread per node reduces its 8 children; level 1 is widened by one ulp so that
// every node box is STRICTLY larger than the voxel boxes it encloses (keeps the NaN cases conservative)
__global__ void aabb_build_up_kernel(float* __restrict__ box_all, long long tree_stride_box, int total,
                                     int off_child, int cnt_child, int off_parent, int cnt_parent, int widen) {
  float* box = box_all + (long long)blockIdx.y * tree_stride_box;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int pad = (cnt_parent + 31) / 32 * 32;
  if (j >= pad) return;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (j < cnt_parent) {
    for (int c = j * 8; c < min(j * 8 + 8, cnt_child); ++c) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(lo[a], box[(long long)a * total + off_child + c]);
        hi[a] = fmaxf(hi[a], box[(long long)(3 + a) * total + off_child + c]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float l = j < cnt_parent ? lo[a] : 0.0f, h = j < cnt_parent ? hi[a] : 0.0f;
    if (widen && j < cnt_parent) { l = nextafterf(l, -INFINITY); h = nextafterf(h, INFINITY); }
    box[(long long)a * total + off_parent + j] = l;
    box[(long long)(3 + a) * total + off_parent + j] = h;
  }
}

// ---- traversal ------------------------------------------------------------------------------------
extern __shared__ __align__(128) float aabb_smem[];   // [6][sm_nodes] staged boxes, then per-warp lists + hit buffers

enum AabbMode { kModeIndexOrder = 0, kModeDepthSorted = 1, kModeAnyHit = 2 };

struct AabbRay {
  float ox, oy, oz, ix, iy, iz;
  bool regular;
};

struct AabbWarpState {
  const float* gbox;      // global SoA of this tree
  int total, sm_nodes, stage_from;
  int* h_idx;             // per-warp hit buffers (shared memory)
  float* h_min;
  float* h_max;
  int n_max, list_cap;
};

// One level of the breadth-first walk: the ascending list `cur` (n_cur hit nodes of level L+1) is consumed four
// nodes per step, their 4 x 8 children (level L) are tested one per lane, and the hit children are appended in
// ascending order to `nxt` (L > 0) or emitted as hits (L == 0).  LEAF / STAGED / REGULAR are hoisted out of the
// loop as template parameters; invalid lanes load a clamped index instead of branching.
//   lo_node : children below this index hold only leaves that were already handled (continuation passes)
//   returns the number of entries written to `nxt` (list_cap + 1 if it overflowed; next_lo then holds the first
//   leaf of the first dropped child)
template <bool LEAF, bool STAGED, bool REGULAR, int MODE>
__device__ __forceinline__ int aabb_level(const AabbWarpState& w, const AabbRay& r, int offL, int cntL, int shift,
                                          int lo_node, const int* cur, int n_cur, int* nxt, int& cnt, int limit,
                                          long long& next_lo) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7, quad = lane >> 3;
  const unsigned lt = (1u << lane) - 1u;
  int n_nxt = 0;
  for (int base = 0; base < n_cur && cnt < limit && n_nxt <= w.list_cap; base += 4) {
    const int p = base + quad;
    const int c = cur[min(p, n_cur - 1)] * 8 + sub;
    const bool valid = p < n_cur && c < cntL && c >= lo_node;
    const int pos = offL + min(c, cntL - 1);
    float lx, ly, lz, hx, hy, hz;
    if (STAGED) {
      const float* s = aabb_smem + (pos - w.stage_from);
      const int st = w.sm_nodes;
      lx = s[0]; ly = s[st]; lz = s[2 * st]; hx = s[3 * st]; hy = s[4 * st]; hz = s[5 * st];
    } else {
      const float* g = w.gbox + pos;
      const int st = w.total;
      lx = __ldg(g); ly = __ldg(g + st); lz = __ldg(g + 2 * st);
      hx = __ldg(g + 3 * (long long)st); hy = __ldg(g + 4 * (long long)st); hz = __ldg(g + 5 * (long long)st);
    }
    float tn, tf;
    bool hit;
    if (REGULAR) {   // no NaN possible: fmin/fmax ordering == the reference's swap, at leaves and nodes alike
      const float a0 = __fmul_rn(__fsub_rn(lx, r.ox), r.ix), b0 = __fmul_rn(__fsub_rn(hx, r.ox), r.ix);
      const float a1 = __fmul_rn(__fsub_rn(ly, r.oy), r.iy), b1 = __fmul_rn(__fsub_rn(hy, r.oy), r.iy);
      const float a2 = __fmul_rn(__fsub_rn(lz, r.oz), r.iz), b2 = __fmul_rn(__fsub_rn(hz, r.oz), r.iz);
      tn = fmaxf(fmaxf(0.0f, fminf(a0, b0)), fmaxf(fminf(a1, b1), fminf(a2, b2)));
      tf = fminf(fminf(100000.0f, fmaxf(a0, b0)), fminf(fmaxf(a1, b1), fmaxf(a2, b2)));
      hit = tn <= tf;
    } else if (LEAF) {
      hit = slab_exact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz, tn, tf);
    } else {
      hit = slab_enclosing(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz);
    }
    hit = hit && valid;
    const unsigned m = __ballot_sync(NSVF_FULL_MASK, hit);
    if (m == 0u) continue;
    const int rank = __popc(m & lt);
    if (LEAF) {
      if (MODE != kModeAnyHit && hit && cnt + rank < w.n_max) {
        w.h_idx[cnt + rank] = c;
        w.h_min[cnt + rank] = tn;
        w.h_max[cnt + rank] = tf;
      }
      cnt += __popc(m);
    } else {
      const int slot = n_nxt + rank;
      if (hit && slot < w.list_cap) nxt[slot] = c;
      const int tot = n_nxt + __popc(m);
      if (tot > w.list_cap) {   // overflow: remember the first leaf of the first dropped child (ascending order)
        const unsigned owner = __ballot_sync(NSVF_FULL_MASK, hit && slot == w.list_cap);
        const int c_first = __shfl_sync(NSVF_FULL_MASK, c, __ffs(owner) - 1);
        next_lo = ((long long)c_first) << shift;
        n_nxt = w.list_cap + 1;
      } else {
        n_nxt = tot;
      }
    }
  }
  return n_nxt;
}

template <bool LEAF, int MODE>
__device__ __forceinline__ int aabb_level_dispatch(const AabbWarpState& w, const AabbRay& r, bool staged, int offL,
                                                   int cntL, int shift, int lo_node, const int* cur, int n_cur,
                                                   int* nxt, int& cnt, int limit, long long& next_lo) {
  if (r.regular) {
    if (staged) return aabb_level<LEAF, true, true, MODE>(w, r, offL, cntL, shift, lo_node, cur, n_cur, nxt, cnt, limit, next_lo);
    return aabb_level<LEAF, false, true, MODE>(w, r, offL, cntL, shift, lo_node, cur, n_cur, nxt, cnt, limit, next_lo);
  }
  if (staged) return aabb_level<LEAF, true, false, MODE>(w, r, offL, cntL, shift, lo_node, cur, n_cur, nxt, cnt, limit, next_lo);
  return aabb_level<LEAF, false, false, MODE>(w, r, offL, cntL, shift, lo_node, cur, n_cur, nxt, cnt, limit, next_lo);
}

// Warp-level stable sort of the first `cnt` hits by entry depth (ties keep ascending slot = ascending voxel
// index).  Bitonic network held in REGISTERS: element t = r*32 + lane, R elements per lane; partners at distance
// < 32 are exchanged with __shfl_xor_sync, partners at distance >= 32 live in the same lane.  perm[] (shared
// memory, >= cnt ints) receives the source slot of every sorted position.
template <int R>
__device__ __forceinline__ void aabb_bitonic_regs(const float* h_min, int cnt, int* perm) {
  const int lane = threadIdx.x & 31;
  float d[R];
  int sl[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = r * 32 + lane;
    sl[r] = t;
    d[r] = t < cnt ? h_min[t] : INFINITY;
  }
#pragma unroll
  for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {   // partner in the same lane: r ^ (j / 32)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int q = r ^ (j >> 5);
          if (q > r) {
            const int t = r * 32 + lane;
            const bool up = (t & k) == 0;
            const bool gt = (d[r] > d[q]) || (d[r] == d[q] && sl[r] > sl[q]);
            if (gt == up) {
              const float td = d[r]; d[r] = d[q]; d[q] = td;
              const int ts = sl[r]; sl[r] = sl[q]; sl[q] = ts;
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int t = r * 32 + lane;
          const float od = __shfl_xor_sync(NSVF_FULL_MASK, d[r], j);
          const int os = __shfl_xor_sync(NSVF_FULL_MASK, sl[r], j);
          const bool up = (t & k) == 0;
          const bool lower = (lane & j) == 0;
          const bool mine_gt = (d[r] > od) || (d[r] == od && sl[r] > os);
          // the lower position of the pair keeps the smaller key when sorting up, the larger when sorting down
          const bool take_other = (up == lower) ? mine_gt : !mine_gt;
          if (take_other) { d[r] = od; sl[r] = os; }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = r * 32 + lane;
    if (t < cnt) perm[t] = sl[r];
  }
}

__device__ __forceinline__ void aabb_sort_by_depth(const float* h_min, int cnt, int* perm) {
  if (cnt <= 16) {   // rank sort: one element per lane, keys broadcast from shared memory (cnt iterations)
    const int lane = threadIdx.x & 31;
    if (lane < cnt) {
      const float d = h_min[lane];
      int rank = 0;
      for (int j = 0; j < cnt; ++j) {
        const float e = h_min[j];
        rank += (e < d) || (e == d && j < lane);
      }
      perm[rank] = lane;
    }
  } else if (cnt <= 32) aabb_bitonic_regs<1>(h_min, cnt, perm);
  else if (cnt <= 64) aabb_bitonic_regs<2>(h_min, cnt, perm);
  else if (cnt <= 128) aabb_bitonic_regs<4>(h_min, cnt, perm);
  else if (cnt <= 256) aabb_bitonic_regs<8>(h_min, cnt, perm);
  else {             // very long hit lists: bitonic network in shared memory over (depth, slot) keys
    const int lane = threadIdx.x & 31;
    int n2 = 512;
    while (n2 < cnt) n2 <<= 1;
    for (int t = lane; t < n2; t += 32) perm[t] = t;
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < n2; t += 32) {
          const int u = t ^ j;
          if (u > t) {
            const int pa = perm[t], pb = perm[u];
            const float da = pa < cnt ? h_min[pa] : INFINITY, db = pb < cnt ? h_min[pb] : INFINITY;
            const bool a_gt_b = (da > db) || (da == db && pa > pb);
            const bool up = (t & k) == 0;
            if (a_gt_b == up) { perm[t] = pb; perm[u] = pa; }
          }
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(kAabbWarps * 32)
aabb_intersect_kernel(const __grid_constant__ AabbTree tree, long long tree_stride_box, long long rays_per_tree,
                      int n_max, int sort_slots, int list_cap, float empty_depth,
                      const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                      int* __restrict__ out_idx, float* __restrict__ out_min, float* __restrict__ out_max,
                      unsigned char* __restrict__ out_hit, const unsigned char* __restrict__ grid_ws,
                      size_t grid_set_bytes) {
  __shared__ __align__(8) uint64_t bar;
  if (voxel_grid_usable(grid_ws, grid_set_bytes, blockIdx.y)) return;   // the lattice walk (voxel_grid.cu) has this set
  const int sm_nodes = tree.total - tree.stage_from;   // multiple of 32
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
        tma_bulk_g2s(aabb_smem + (size_t)a * sm_nodes, gbox + (long long)a * tree.total + tree.stage_from, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  }

  const int hit_words = MODE == kModeAnyHit ? 0 : 3 * n_max + sort_slots;
  const int per_warp = 2 * list_cap + hit_words;
  float* wbase = aabb_smem + (size_t)6 * sm_nodes + (size_t)warp * per_warp;
  int* list_a = reinterpret_cast<int*>(wbase);
  int* list_b = list_a + list_cap;
  float* h = wbase + 2 * list_cap;
  int* h_idx = reinterpret_cast<int*>(h);
  float* h_min = h + n_max;
  float* h_max = h + 2 * n_max;
  int* perm = reinterpret_cast<int*>(h + 3 * n_max);
  const int top = tree.nlevels - 1;
  const int limit = MODE == kModeAnyHit ? 1 : n_max;
  AabbWarpState w;
  w.gbox = gbox; w.total = tree.total; w.sm_nodes = sm_nodes; w.stage_from = tree.stage_from;
  w.h_idx = h_idx; w.h_min = h_min; w.h_max = h_max; w.n_max = n_max;
  // any-hit: keep only one 4-node batch per level, i.e. walk depth-first by batches (the continuation pass resumes
  // behind the explored subtree), so a ray stops at its first hit instead of finishing whole levels
  if (MODE == kModeAnyHit) list_cap = 4;
  w.list_cap = list_cap;

  const long long ray_base = (long long)blockIdx.y * rays_per_tree;
  for (long long rr = (long long)blockIdx.x * kAabbWarps + warp; rr < rays_per_tree;
       rr += (long long)gridDim.x * kAabbWarps) {
    const long long ray = ray_base + rr;
    float v = 0.f;   // lanes 0..2 load the origin, 3..5 the direction; broadcast
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

    int cnt = 0;
    long long leaf_lo = 0;          // leaves below this index are already done (continuation passes)
    for (;;) {
      long long next_lo = -1;       // >= 0: some level's list overflowed; leaves from here on need another pass
      // top level: <= 32 nodes, one per lane; its hits seed list_a (as "parents" of a virtual level: we store the
      // node ids themselves and let the level routine expand children, so the top is handled by a direct test)
      int n_cur;
      {
        const int c = lane;
        const bool valid = c < tree.cnt[top] && c >= (int)(leaf_lo >> (3 * top));
        const int pos = tree.off[top] + min(c, tree.cnt[top] - 1);
        const bool staged = tree.off[top] >= tree.stage_from;
        float lx, ly, lz, hx, hy, hz;
        if (staged) {
          const float* sp = aabb_smem + (pos - tree.stage_from);
          lx = sp[0]; ly = sp[sm_nodes]; lz = sp[2 * sm_nodes]; hx = sp[3 * sm_nodes]; hy = sp[4 * sm_nodes]; hz = sp[5 * sm_nodes];
        } else {
          const float* g = gbox + pos;
          const long long st = tree.total;
          lx = g[0]; ly = g[st]; lz = g[2 * st]; hx = g[3 * st]; hy = g[4 * st]; hz = g[5 * st];
        }
        float tn = 0.f, tf = 0.f;
        bool hit;
        if (top == 0) hit = slab_exact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz, tn, tf);
        else if (r.regular) hit = slab_sorted(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, r.ix < 0.f ? hx : lx, r.iy < 0.f ? hy : ly,
                                              r.iz < 0.f ? hz : lz, r.ix < 0.f ? lx : hx, r.iy < 0.f ? ly : hy,
                                              r.iz < 0.f ? lz : hz, tn, tf);
        else hit = slab_enclosing(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz);
        hit = hit && valid;
        const unsigned m = __ballot_sync(NSVF_FULL_MASK, hit);
        const int rank = __popc(m & ((1u << lane) - 1u));
        if (top == 0) {
          if (MODE != kModeAnyHit && hit && cnt + rank < n_max) { h_idx[cnt + rank] = c; h_min[cnt + rank] = tn; h_max[cnt + rank] = tf; }
          cnt += __popc(m);
          n_cur = 0;
        } else {
          if (hit) list_a[rank] = c;
          n_cur = __popc(m);
        }
      }
      __syncwarp();
      int* cur = list_a;
      int* nxt = list_b;
      for (int L = top - 1; L >= 1 && n_cur > 0; --L) {
        const int n_nxt = aabb_level_dispatch<false, MODE>(w, r, tree.off[L] >= tree.stage_from, tree.off[L], tree.cnt[L],
                                                           3 * L, (int)(leaf_lo >> (3 * L)), cur, n_cur, nxt, cnt, limit,
                                                           next_lo);
        __syncwarp();
        n_cur = min(n_nxt, w.list_cap);
        int* t = cur; cur = nxt; nxt = t;
      }
      if (top >= 1 && n_cur > 0 && cnt < limit) {
        aabb_level_dispatch<true, MODE>(w, r, tree.off[0] >= tree.stage_from, tree.off[0], tree.cnt[0], 0,
                                        (int)leaf_lo, cur, n_cur, nxt, cnt, limit, next_lo);
        __syncwarp();
      }
      // a truncated list only ever dropped nodes whose leaves come AFTER every leaf handled so far, so the
      // collected hits are exactly the first ones; continue from the first dropped leaf if more are needed
      if (next_lo < 0 || cnt >= limit) break;
      leaf_lo = next_lo;
    }

    if constexpr (MODE == kModeAnyHit) {
      if (lane == 0) out_hit[ray] = cnt > 0;
    } else {
      cnt = min(cnt, n_max);
      __syncwarp();
      const long long row = ray * n_max;
      if constexpr (MODE == kModeDepthSorted) {
        aabb_sort_by_depth(h_min, cnt, perm);
        for (int l = lane; l < n_max; l += 32) {
          const bool ok = l < cnt;
          const int src = ok ? perm[l] : 0;
          out_idx[row + l] = ok ? h_idx[src] : -1;
          out_min[row + l] = ok ? h_min[src] : empty_depth;
          out_max[row + l] = ok ? h_max[src] : empty_depth;
        }
        if (out_hit != nullptr && lane == 0) out_hit[ray] = cnt > 0;
      } else {
        for (int l = lane; l < n_max; l += 32) {
          const bool ok = l < cnt;
          out_idx[row + l] = ok ? h_idx[l] : -1;
          out_min[row + l] = ok ? h_min[l] : empty_depth;
          out_max[row + l] = ok ? h_max[l] : empty_depth;
        }
      }
      __syncwarp();
    }
  }
}

// ---- small voxel sets: one thread per ray over ALL voxels, boxes broadcast from shared memory ---------------------
// For a few hundred voxels (the first phase of training: 343-512 voxels) no hierarchy pays for itself: consecutive
// index groups straddle grid rows, so almost every node is hit.  The reference's linear scan is the right
// algorithm there; this is that scan with everything around it fixed: exact voxel boxes (c -+ hv) staged once
// per CTA as 32-byte records, every lane of a warp reads the SAME box (two LDS.128 broadcasts, no conflicts), 1/dir
// hoisted out of the loop, the branch-free min/max form of the slab test for regular rays (~17 instructions per
// test instead of ~40), rows pre-filled with coalesced stores by the whole CTA.
constexpr int kSmallThreads = 256;
constexpr int kSmallMaxVoxels = 1024;

template <int MODE>
__global__ void __launch_bounds__(kSmallThreads)
aabb_small_kernel(const __grid_constant__ AabbTree tree, long long tree_stride_box, long long rays_per_tree, int n,
                  int n_max, const float* __restrict__ ray_start, const float* __restrict__ ray_dir,
                  int* __restrict__ out_idx, float* __restrict__ out_min, float* __restrict__ out_max,
                  unsigned char* __restrict__ out_hit, const unsigned char* __restrict__ grid_ws,
                  size_t grid_set_bytes) {
  if (voxel_grid_usable(grid_ws, grid_set_bytes, blockIdx.y)) return;   // the lattice walk (voxel_grid.cu) has this set
  // stage the exact voxel boxes as 32-byte records {lo.xyz, hi.x | hi.yz, -, -}: one test = two LDS.128 broadcasts;
  // behind them the level-1 boxes (union of 8 consecutive voxels, widened by an ulp like every internal node): a ray
  // first tests the group and scans its 8 voxels only when the group is hit — ~2.4x fewer box tests at 343 voxels
  // (any-hit over 2.56 M rays: 0.41 -> 0.14 ms), hits still emitted in ascending voxel index.
  // Measured and rejected for the index-order mode: recording hits as a bit mask and expanding it warp-cooperatively into
  // coalesced rows (no pre-fill, no scattered stores) — 1.45 ms against 1.30 ms for the scattered stores below; the
  // expansion's ~170 warp instructions per ray cost more than the partial-sector stores they remove.
  const float* gbox = tree.box + (long long)blockIdx.y * tree_stride_box;
  float4* sbox = reinterpret_cast<float4*>(aabb_smem);
  const long long st = tree.total;
  for (int k = threadIdx.x; k < n; k += kSmallThreads) {
    const float* g = gbox + k;
    sbox[2 * k] = make_float4(g[0], g[st], g[2 * st], g[3 * st]);
    sbox[2 * k + 1] = make_float4(g[4 * st], g[5 * st], 0.f, 0.f);
  }
  const int n1 = tree.nlevels > 1 ? tree.cnt[1] : 0;
  float4* sgrp = sbox + 2 * n;
  for (int k = threadIdx.x; k < n1; k += kSmallThreads) {
    const float* g = gbox + tree.off[1] + k;
    sgrp[2 * k] = make_float4(g[0], g[st], g[2 * st], g[3 * st]);
    sgrp[2 * k + 1] = make_float4(g[4 * st], g[5 * st], 0.f, 0.f);
  }
  __syncthreads();
  const long long ray_base = (long long)blockIdx.y * rays_per_tree;

  for (long long tile = (long long)blockIdx.x * kSmallThreads; tile < rays_per_tree;
       tile += (long long)gridDim.x * kSmallThreads) {
    const long long tile_rays = min((long long)kSmallThreads, rays_per_tree - tile);
    if constexpr (MODE != kModeAnyHit) {   // coalesced pre-fill of this tile's rows
      const long long base = (ray_base + tile) * n_max, cells = tile_rays * n_max;
      for (long long c = threadIdx.x; c < cells; c += kSmallThreads) {
        out_idx[base + c] = -1;
        out_min[base + c] = 0.0f;
        out_max[base + c] = 0.0f;
      }
      __syncthreads();
    }
    if (threadIdx.x < tile_rays) {
      const long long ray = ray_base + tile + threadIdx.x;
      const float ox = ray_start[ray * 3 + 0], oy = ray_start[ray * 3 + 1], oz = ray_start[ray * 3 + 2];
      const float ix = ref_rcp(ray_dir[ray * 3 + 0]), iy = ref_rcp(ray_dir[ray * 3 + 1]),
                  iz = ref_rcp(ray_dir[ray * 3 + 2]);
      const bool regular = regular_component(ox, ix) && regular_component(oy, iy) && regular_component(oz, iz);
      const long long row = ray * n_max;
      int cnt = 0;
      const int limit = MODE == kModeAnyHit ? 1 : n_max;
      for (int k = 0; k < n && cnt < limit; ++k) {
        if (n1 > 0 && (k & 7) == 0) {     // entering a new group of 8: skip it when the ray misses its enclosing box
          const float4 g0 = sgrp[2 * (k >> 3)], g1 = sgrp[2 * (k >> 3) + 1];
          if (!slab_enclosing(ox, oy, oz, ix, iy, iz, g0.x, g0.y, g0.z, g0.w, g1.x, g1.y)) { k += 7; continue; }
        }
        float tn, tf;
        bool hit;
        const float4 q0 = sbox[2 * k], q1 = sbox[2 * k + 1];
        if (regular) {
          const float a0 = __fmul_rn(__fsub_rn(q0.x, ox), ix), b0 = __fmul_rn(__fsub_rn(q0.w, ox), ix);
          const float a1 = __fmul_rn(__fsub_rn(q0.y, oy), iy), b1 = __fmul_rn(__fsub_rn(q1.x, oy), iy);
          const float a2 = __fmul_rn(__fsub_rn(q0.z, oz), iz), b2 = __fmul_rn(__fsub_rn(q1.y, oz), iz);
          tn = fmaxf(fmaxf(0.0f, fminf(a0, b0)), fmaxf(fminf(a1, b1), fminf(a2, b2)));
          tf = fminf(fminf(100000.0f, fmaxf(a0, b0)), fminf(fmaxf(a1, b1), fmaxf(a2, b2)));
          hit = tn <= tf;
        } else {
          hit = slab_exact(ox, oy, oz, ix, iy, iz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn, tf);
        }
        if (hit) {
          if constexpr (MODE != kModeAnyHit) {
            out_idx[row + cnt] = k;
            out_min[row + cnt] = tn;
            out_max[row + cnt] = tf;
          }
          ++cnt;
        }
      }
      if constexpr (MODE == kModeAnyHit) out_hit[ray] = cnt > 0;
    }
    if constexpr (MODE != kModeAnyHit) __syncthreads();
  }
}

// Stand-alone epilogue for hit lists produced elsewhere (the octree traversal): in-place masked_fill + sort by entry
// depth + any(), i.e. SparseVoxelEncoder.ray_intersect's post-processing (encoder.py:519-524), one warp per ray.
__global__ void __launch_bounds__(kAabbWarps * 32)
sort_hits_kernel(long long rays, int n_max, int sort_slots, float empty_depth, int* __restrict__ idx,
                 float* __restrict__ dmin, float* __restrict__ dmax, unsigned char* __restrict__ out_hit,
                 const int* __restrict__ walk_active, long long rays_per_tree, const unsigned char* __restrict__ defer) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* h = aabb_smem + (size_t)warp * (3 * n_max + sort_slots);
  int* h_idx = reinterpret_cast<int*>(h);
  float* h_min = h + n_max;
  float* h_max = h + 2 * n_max;
  int* perm = reinterpret_cast<int*>(h + 3 * n_max);
  for (long long ray = (long long)blockIdx.x * kAabbWarps + warp; ray < rays; ray += (long long)gridDim.x * kAabbWarps) {
    // octree queries answered by the lattice walk: only the rays it deferred to the traversal need this pass
    if (walk_active != nullptr && walk_active[ray / rays_per_tree] != 0 && defer[ray] == 0) continue;
    const long long row = ray * n_max;
    // compact the valid entries to the front, keeping their order (the reference sorts all slots; -1 slots carry
    // MAX_DEPTH and end up last)
    int cnt = 0;
    for (int l0 = 0; l0 < n_max; l0 += 32) {
      const int l = l0 + lane;
      const int v = l < n_max ? idx[row + l] : -1;
      const unsigned m = __ballot_sync(NSVF_FULL_MASK, v != -1);
      if (v != -1) {
        const int r = cnt + __popc(m & ((1u << lane) - 1u));
        h_idx[r] = v;
        h_min[r] = dmin[row + l];
        h_max[r] = dmax[row + l];
      }
      cnt += __popc(m);
    }
    __syncwarp();
    aabb_sort_by_depth(h_min, cnt, perm);
    for (int l = lane; l < n_max; l += 32) {
      const bool ok = l < cnt;
      const int src = ok ? perm[l] : 0;
      idx[row + l] = ok ? h_idx[src] : -1;
      dmin[row + l] = ok ? h_min[src] : empty_depth;
      dmax[row + l] = ok ? h_max[src] : empty_depth;
    }
    if (out_hit != nullptr && lane == 0) out_hit[ray] = cnt > 0;
    __syncwarp();
  }
}

static size_t aabb_tree_floats(int n) { return (size_t)6 * aabb_layout(n).total; }

template <int MODE>
static int aabb_launch(cudaStream_t stream, dim3 grid, size_t smem, const AabbTree& tree, long long tree_stride_box,
                       long long rays_per_tree, int n_max, int sort_slots, int list_cap, float empty_depth,
                       const float* ray_start, const float* ray_dir, int* idx, float* dmin, float* dmax,
                       unsigned char* hit, const unsigned char* grid_ws, size_t grid_set_bytes) {
  NSVF_CUDA_OK(cudaFuncSetAttribute(aabb_intersect_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
  // with a lattice workspace this launch is the stand-by of the walk (it returns at once when the lattice is usable):
  // it is timed under its own name so that the live kernel timers see the launch that does the work
  const char* kname = grid_ws != nullptr ? "aabb_hierarchy_standby_kernel"
                      : MODE == kModeAnyHit ? "aabb_hit_mask_kernel"
                      : (MODE == kModeDepthSorted ? "aabb_intersect_sorted_kernel" : "aabb_intersect_kernel");
  NSVF_TIMED_LAUNCH(kname, stream,
                    (aabb_intersect_kernel<MODE><<<grid, kAabbWarps * 32, smem, stream>>>(
                        tree, tree_stride_box, rays_per_tree, n_max, sort_slots, list_cap, empty_depth, ray_start,
                        ray_dir, idx, dmin, dmax, hit, grid_ws, grid_set_bytes)));
  return 0;
}

}  // namespace nsvf

using namespace nsvf;

static size_t aabb_tree_bytes(int n, int n_trees) {   // all hierarchies, rounded so that the lattices behind stay aligned
  return (aabb_tree_floats(n) * sizeof(float) * (size_t)n_trees + 127) / 128 * 128;
}

extern "C" size_t nsvf_aabb_workspace_bytes(int n, int n_trees) {
  if (n <= 0 || n_trees <= 0) return 0;
  return aabb_tree_bytes(n, n_trees) + voxel_grid_bytes(n) * (size_t)n_trees;
}

// mode: 0 = reference order (ascending voxel index), 1 = sorted by entry depth, 2 = any-hit mask only
__global__ void fill_f32_kernel(float* __restrict__ a, float* __restrict__ b, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    a[i] = v;
    b[i] = v;
  }
}

// phase: kBuild fills the workspace (hierarchy + lattice of every voxel set), kTraverse intersects rays with a filled one
enum { kBuild = 1, kTraverse = 2 };

static int aabb_run(cudaStream_t stream, int phase, int mode, int b, int n, int m, float voxelsize, int n_max,
                    float empty_depth, const float* ray_start, const float* ray_dir, const float* points,
                    long long points_batch_stride, int* idx, float* min_depth, float* max_depth,
                    unsigned char* hit, void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(b >= 0 && n >= 0 && m >= 0 && n_max >= 0, "aabb_intersect: negative size");
  const bool traverse = (phase & kTraverse) != 0;
  if (traverse && (b == 0 || m == 0)) return 0;
  if (traverse && mode != kModeAnyHit && n_max == 0) return 0;
  const long long rays = (long long)b * m;
  if (n == 0 && !traverse) return 0;
  if (n == 0) {  // nothing to hit
    if (mode != kModeAnyHit) {
      NSVF_CUDA_OK(cudaMemsetAsync(idx, 0xff, sizeof(int) * rays * n_max, stream));
      if (empty_depth == 0.0f) {
        NSVF_CUDA_OK(cudaMemsetAsync(min_depth, 0, sizeof(float) * rays * n_max, stream));
        NSVF_CUDA_OK(cudaMemsetAsync(max_depth, 0, sizeof(float) * rays * n_max, stream));
      } else {   // sorted mode of an empty voxel set (everything pruned): every slot holds the fill depth
        const long long cells = rays * n_max;
        fill_f32_kernel<<<(unsigned)((cells + 255) / 256 < 65535 ? (cells + 255) / 256 : 65535), 256, 0, stream>>>(
            min_depth, max_depth, cells, empty_depth);
        NSVF_LAUNCH_OK("fill_f32_kernel");
      }
    }
    if (hit != nullptr) NSVF_CUDA_OK(cudaMemsetAsync(hit, 0, rays, stream));
    return 0;
  }
  NSVF_REQUIRE(points_batch_stride == 0 || points_batch_stride >= (long long)n * 3,
               "aabb_intersect: points_batch_stride must be 0 (shared voxel set) or >= 3*n");
  const int n_trees = points_batch_stride == 0 ? 1 : b;
  const size_t need = nsvf_aabb_workspace_bytes(n, n_trees);
  NSVF_REQUIRE(workspace != nullptr && workspace_bytes >= need,
               "aabb_intersect: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  NSVF_REQUIRE(((uintptr_t)workspace & 127) == 0, "aabb_intersect: workspace must be 128-byte aligned");

  // the voxel lattice and its walk (voxel_grid.cu); when the voxel set is not a lattice the walk returns at once and
  // the hierarchy kernels below do the work — and the other way round
  const size_t grid_set_bytes = voxel_grid_bytes(n);
  unsigned char* grid_ws = grid_set_bytes ? (unsigned char*)workspace + aabb_tree_bytes(n, n_trees) : nullptr;
  // index-order queries stay on the hierarchy, which produces that order natively (the walk finds hits in depth order
  // and would have to sort ~76 indices per ray at 112 k voxels: measured 13.2 ms against 2.3 ms)
  if (phase == (kBuild | kTraverse) && mode == kModeIndexOrder) grid_ws = nullptr;
  if (grid_ws != nullptr && (phase & kBuild) &&
      voxel_grid_build(stream, n_trees, n, points, points_batch_stride, voxelsize, grid_ws, grid_set_bytes))
    return 1;
  if (mode == kModeIndexOrder) grid_ws = nullptr;
  if (grid_ws != nullptr && traverse &&
      voxel_grid_walk(stream, mode, grid_ws, grid_set_bytes, n_trees, n, points, points_batch_stride, voxelsize,
                      n_trees == 1 ? rays : m, n_max, empty_depth, ray_start, ray_dir, idx, min_depth, max_depth, hit))
    return 1;

  AabbLayout L = aabb_layout(n);
  float* box = (float*)workspace;
  const long long tree_stride_box = (long long)6 * L.total;
  const float half_voxel = voxelsize * 0.5f;  // reference: float half_voxel = voxelsize * 0.5 (exact)

  if (phase & kBuild) {  // build the hierarchy (O(n), one small launch per level)
    const int n_pad = (n + 31) / 32 * 32;
    dim3 g0((n_pad + 255) / 256, n_trees);
    aabb_build_leaves_kernel<<<g0, 256, 0, stream>>>(points, points_batch_stride, n, half_voxel, box, tree_stride_box,
                                                     L.total);
    NSVF_LAUNCH_OK("aabb_build_leaves_kernel");
    for (int l = 1; l < L.nlevels; ++l) {
      const int pad = (L.cnt[l] + 31) / 32 * 32;
      dim3 g((pad + 255) / 256, n_trees);
      aabb_build_up_kernel<<<g, 256, 0, stream>>>(box, tree_stride_box, L.total, L.off[l - 1], L.cnt[l - 1], L.off[l],
                                                  L.cnt[l], l == 1);
      NSVF_LAUNCH_OK("aabb_build_up_kernel");
    }
  }

  if (!traverse) return 0;

  AabbTree tree;
  tree.box = box;
  tree.total = L.total;
  tree.nlevels = L.nlevels;
  for (int l = 0; l < kAabbMaxLevels; ++l) { tree.cnt[l] = L.cnt[l]; tree.off[l] = L.off[l]; }
  tree.stage_from = L.stage_from;

  const long long rays_per_tree = n_trees == 1 ? rays : m;
  const int sm_nodes = L.total - L.stage_from;
  int sort_slots = 0;
  if (mode == kModeDepthSorted) {
    sort_slots = 64;
    while (sort_slots < n_max) sort_slots <<= 1;
  }
  int list_cap = kAabbListCap;
  if (const char* e = getenv("NSVF_AABB_LIST_CAP")) list_cap = atoi(e) < 32 ? 32 : atoi(e);   // test hook (>= 32: the top level writes up to 32 entries)
  const int per_warp = 2 * list_cap + (mode == kModeAnyHit ? 0 : 3 * n_max + sort_slots);
  size_t smem = ((size_t)6 * sm_nodes + (size_t)kAabbWarps * per_warp) * sizeof(float);
  NSVF_REQUIRE(smem <= 200 * 1024, "aabb_intersect: n_max=%d needs %zu B of shared memory", n_max, smem);
  int blocks_per_sm = (int)((220 * 1024) / (smem + 1024));
  blocks_per_sm = blocks_per_sm < 1 ? 1 : (blocks_per_sm > 6 ? 6 : blocks_per_sm);
  long long want = (rays_per_tree + kAabbWarps - 1) / kAabbWarps;
  long long cap = (long long)num_sms() * blocks_per_sm;
  if (n_trees > 1) cap = (cap + n_trees - 1) / n_trees;
  int gx = (int)(want < cap ? want : cap);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_trees);
  if (n <= kSmallMaxVoxels && mode != kModeDepthSorted && getenv("NSVF_AABB_NO_SMALL") == nullptr) {
    const size_t sm = (size_t)8 * (n + (n + 7) / 8 + 1) * sizeof(float);
    long long want_s = (rays_per_tree + kSmallThreads - 1) / kSmallThreads, cap_s = (long long)num_sms() * 8;
    if (n_trees > 1) cap_s = (cap_s + n_trees - 1) / n_trees;
    dim3 gs((unsigned)(want_s < cap_s ? want_s : cap_s), n_trees);
    if (mode == kModeAnyHit) {
      NSVF_TIMED_LAUNCH(grid_ws != nullptr ? "aabb_hierarchy_standby_kernel" : "aabb_hit_mask_kernel", stream,
                        (aabb_small_kernel<kModeAnyHit><<<gs, kSmallThreads, sm, stream>>>(
                            tree, tree_stride_box, rays_per_tree, n, 1, ray_start, ray_dir, idx, min_depth, max_depth, hit,
                            grid_ws, grid_set_bytes)));
    } else {
      NSVF_TIMED_LAUNCH("aabb_intersect_kernel", stream,
                        (aabb_small_kernel<kModeIndexOrder><<<gs, kSmallThreads, sm, stream>>>(
                            tree, tree_stride_box, rays_per_tree, n, n_max, ray_start, ray_dir, idx, min_depth, max_depth, hit,
                            grid_ws, grid_set_bytes)));
    }
    return 0;
  }
  const int nm = mode == kModeAnyHit ? 0 : n_max;
  switch (mode) {
    case kModeIndexOrder:
      return aabb_launch<kModeIndexOrder>(stream, grid, smem, tree, tree_stride_box, rays_per_tree, nm, sort_slots,
                                          list_cap, empty_depth, ray_start, ray_dir, idx, min_depth, max_depth, hit, grid_ws,
                                          grid_set_bytes);
    case kModeDepthSorted:
      return aabb_launch<kModeDepthSorted>(stream, grid, smem, tree, tree_stride_box, rays_per_tree, nm, sort_slots,
                                           list_cap, empty_depth, ray_start, ray_dir, idx, min_depth, max_depth, hit, grid_ws,
                                          grid_set_bytes);
    default:
      return aabb_launch<kModeAnyHit>(stream, grid, smem, tree, tree_stride_box, rays_per_tree, nm, sort_slots,
                                      list_cap, empty_depth, ray_start, ray_dir, idx, min_depth, max_depth, hit, grid_ws,
                                          grid_set_bytes);
  }
}

extern "C" int nsvf_aabb_intersect(nsvf_stream_t stream, int b, int n, int m, float voxelsize, int n_max,
                                   const float* ray_start, const float* ray_dir, const float* points,
                                   long long points_batch_stride, int* idx, float* min_depth, float* max_depth,
                                   void* workspace, size_t workspace_bytes) {
  return aabb_run((cudaStream_t)stream, kBuild | kTraverse, kModeIndexOrder, b, n, m, voxelsize, n_max, 0.0f, ray_start, ray_dir, points,
                  points_batch_stride, idx, min_depth, max_depth, nullptr, workspace, workspace_bytes);
}

extern "C" int nsvf_aabb_intersect_sorted(nsvf_stream_t stream, int b, int n, int m, float voxelsize, int n_max,
                                          float empty_depth, const float* ray_start, const float* ray_dir,
                                          const float* points, long long points_batch_stride, int* idx,
                                          float* min_depth, float* max_depth, unsigned char* hits, void* workspace,
                                          size_t workspace_bytes) {
  return aabb_run((cudaStream_t)stream, kBuild | kTraverse, kModeDepthSorted, b, n, m, voxelsize, n_max, empty_depth, ray_start, ray_dir,
                  points, points_batch_stride, idx, min_depth, max_depth, hits, workspace, workspace_bytes);
}

extern "C" int nsvf_aabb_hit_mask(nsvf_stream_t stream, int b, int n, int m, float voxelsize, const float* ray_start,
                                  const float* ray_dir, const float* points, long long points_batch_stride,
                                  unsigned char* hits, void* workspace, size_t workspace_bytes) {
  return aabb_run((cudaStream_t)stream, kBuild | kTraverse, kModeAnyHit, b, n, m, voxelsize, 0, 0.0f, ray_start, ray_dir, points,
                  points_batch_stride, nullptr, nullptr, nullptr, hits, workspace, workspace_bytes);
}

extern "C" int nsvf_aabb_prepare(nsvf_stream_t stream, int n_sets, int n, float voxelsize, const float* points,
                                 long long points_batch_stride, void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(n_sets >= 1 && (points_batch_stride != 0 || n_sets == 1),
               "aabb_prepare: a shared voxel set (points_batch_stride 0) is one set");
  return aabb_run((cudaStream_t)stream, kBuild, kModeIndexOrder, n_sets, n, 0, voxelsize, 0, 0.0f, nullptr, nullptr,
                  points, points_batch_stride, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes);
}

extern "C" int nsvf_aabb_intersect_prepared(nsvf_stream_t stream, int mode, int b, int n, int m, float voxelsize,
                                            int n_max, float empty_depth, const float* ray_start,
                                            const float* ray_dir, const float* points, long long points_batch_stride,
                                            int* idx, float* min_depth, float* max_depth, unsigned char* hits,
                                            const void* workspace, size_t workspace_bytes) {
  NSVF_REQUIRE(mode >= 0 && mode <= 2, "aabb_intersect_prepared: mode must be 0 (index order), 1 (sorted) or 2 (any hit)");
  return aabb_run((cudaStream_t)stream, kTraverse, mode, b, n, m, voxelsize, mode == kModeAnyHit ? 0 : n_max,
                  mode == kModeDepthSorted ? empty_depth : 0.0f, ray_start, ray_dir, points, points_batch_stride, idx,
                  min_depth, max_depth, hits, const_cast<void*>(workspace), workspace_bytes);
}

namespace nsvf {
int sort_hits_run(cudaStream_t stream, long long rays, int n_max, float empty_depth, int* idx, float* min_depth,
                  float* max_depth, unsigned char* hits, const int* walk_active, long long rays_per_tree,
                  const unsigned char* defer) {
  NSVF_REQUIRE(rays >= 0 && n_max >= 0, "sort_hits_by_depth: negative size");
  if (rays == 0 || n_max == 0) return 0;
  int sort_slots = 64;
  while (sort_slots < n_max) sort_slots <<= 1;
  const size_t smem = (size_t)kAabbWarps * (3 * n_max + sort_slots) * sizeof(float);
  NSVF_REQUIRE(smem <= 200 * 1024, "sort_hits_by_depth: n_max=%d needs %zu B of shared memory", n_max, smem);
  NSVF_CUDA_OK(cudaFuncSetAttribute(sort_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long want = (rays + kAabbWarps - 1) / kAabbWarps, cap = (long long)num_sms() * 6;
  sort_hits_kernel<<<(unsigned)(want < cap ? want : cap), kAabbWarps * 32, smem, stream>>>(
      rays, n_max, sort_slots, empty_depth, idx, min_depth, max_depth, hits, walk_active,
      rays_per_tree > 0 ? rays_per_tree : 1, defer);
  NSVF_LAUNCH_OK("sort_hits_kernel");
  return 0;
}
}  // namespace nsvf

extern "C" int nsvf_sort_hits_by_depth(nsvf_stream_t stream_, long long rays, int n_max, float empty_depth, int* idx,
                                       float* min_depth, float* max_depth, unsigned char* hits) {
  return sort_hits_run((cudaStream_t)stream_, rays, n_max, empty_depth, idx, min_depth, max_depth, hits, nullptr, 1, nullptr);
}
