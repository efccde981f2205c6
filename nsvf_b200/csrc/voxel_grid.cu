// Ray / voxel-AABB intersection by walking a dense lattice of voxel indices (sm_100a).
//
// Same contract as aabb_intersect.cu (fairnr/clib/src/intersect_gpu.cu:125-167 + the sort / fill / any() of
// fairnr/modules/encoder.py:519-524): the hit decision and the depths of every reported voxel come from the
// reference's slab test on that voxel's own centre (common.cuh), so results are bit-identical.  What changes is which
// voxels get tested.  NSVF's voxel centres always lie on a regular lattice (offset + integer * voxel_size, halved by
// every split), so instead of culling ALL voxels with a hierarchy (≈ 1800 warp instructions per ray at 112 k voxels) a
// ray only visits the lattice cells it passes through — a few hundred 4-byte lookups — and runs the exact test on the
// occupied ones.  Visiting in travel order also yields the hits already sorted by entry depth.
//
//   * voxel_grid_build: min centre -> lattice extent -> clear -> scatter, four small launches, all decisions on the
//     device.  Centres that are not on a common lattice (tolerance 1e-3 cells), duplicate cells, non-finite centres
//     or an extent beyond the cell capacity (16 cells per voxel) set header flags and the hierarchy kernels run
//     instead — the caller never synchronises to find out.
//   * grid_walk_kernel: one thread per ray, layers of the dominant axis in travel order; per layer the cells covered by
//     the ray's span in the two minor axes, widened by `eps` cells (rounding of the reference's own test, lattice
//     tolerance), are CANDIDATES — a superset of the reference's hits; the exact test decides.  Hits are appended to
//     the ray's row as they come; a (depth, index) order violation (grazing hits, ties) is noticed on the fly and
//     repaired by an insertion sort of that row; rows that overflow n_max keep the n_max smallest voxel indices exactly
//     like the reference's index-order scan.  Serves the depth-sorted and the any-hit query; the plain index-order
//     query stays on the hierarchy (it would need a full sort of every row by index).  Rays the walk cannot trust (non-finite, zero or extreme direction, origin
//     more than 1e5 cells away) scan all voxels like the reference does.
//   * the unused tail of 32 consecutive rows (-1 / fill depth) is written cooperatively by the warp (coalesced).
#include <cstdlib>
#include <cstring>

#include "voxel_grid.cuh"

namespace nsvf {

constexpr int kGridThreads = 256;
constexpr int kWalkThreads = 128;
constexpr float kLatticeTol = 1.0e-3f;   // |centre - lattice node| in cells
constexpr int kMaxCellsPerAxis = 1 << 20;

size_t voxel_grid_bytes(int n) {
  if (n <= 0 || getenv("NSVF_AABB_NO_GRID") != nullptr) return 0;
  long long cap = (long long)16 * n;
  if (cap < 4096) cap = 4096;
  if (cap > (1ll << 28)) return 0;   // 1 GiB of cells: keep the hierarchy for such sets
  return (size_t)(sizeof(VoxelGridHeader) + cap * 4 + 127) / 128 * 128;
}

__device__ __forceinline__ VoxelGridHeader* grid_header(unsigned char* ws, size_t per_set_bytes, int set) {
  return reinterpret_cast<VoxelGridHeader*>(ws + (size_t)set * per_set_bytes);
}

__global__ void grid_init_kernel(unsigned char* ws, size_t per_set_bytes, int n_sets, long long cap) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sets) return;
  VoxelGridHeader* h = grid_header(ws, per_set_bytes, s);
  h->min_key[0] = h->min_key[1] = h->min_key[2] = 0x7fffffff;
  h->dims[0] = h->dims[1] = h->dims[2] = 0;
  h->bad = 0;
  h->ok = 0;
  h->cap = cap;
}

__global__ void __launch_bounds__(kGridThreads)
grid_min_kernel(unsigned char* ws, size_t per_set_bytes, const float* __restrict__ points, long long points_stride,
                int n, const int* __restrict__ filter, long long filter_stride) {
  __shared__ int s_key[3];
  VoxelGridHeader* h = grid_header(ws, per_set_bytes, blockIdx.y);
  const float* pts = points + (long long)blockIdx.y * points_stride;
  if (threadIdx.x < 3) s_key[threadIdx.x] = 0x7fffffff;
  __syncthreads();
  const int* keep = filter != nullptr ? filter + (long long)blockIdx.y * filter_stride : nullptr;
  float mn[3] = {INFINITY, INFINITY, INFINITY};
  bool bad = false;
  for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < n; i += gridDim.x * kGridThreads) {
    if (keep != nullptr && keep[i] < 0) continue;      // not a member of this voxel set (octree: not a reachable leaf)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = pts[(long long)i * 3 + a];
      bad = bad || !(fabsf(p) <= 3.0e38f);
      mn[a] = fminf(mn[a], p);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int k = __reduce_min_sync(NSVF_FULL_MASK, float_order_key(mn[a]));
    if ((threadIdx.x & 31) == 0) atomicMin(&s_key[a], k);
  }
  if (__any_sync(NSVF_FULL_MASK, bad) && (threadIdx.x & 31) == 0) atomicOr(&h->bad, 1);
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(&h->min_key[threadIdx.x], s_key[threadIdx.x]);
}

// lattice coordinates of a centre; false when it is off the lattice
__device__ __forceinline__ bool lattice_coords(const float* pts, int i, float gx, float gy, float gz, float voxelsize,
                                               int& qx, int& qy, int& qz) {
  const float rx = __fsub_rn(pts[(long long)i * 3 + 0], gx), ry = __fsub_rn(pts[(long long)i * 3 + 1], gy),
              rz = __fsub_rn(pts[(long long)i * 3 + 2], gz);
  const float fx = rintf(__fdiv_rn(rx, voxelsize)), fy = rintf(__fdiv_rn(ry, voxelsize)),
              fz = rintf(__fdiv_rn(rz, voxelsize));
  const float tol = kLatticeTol * voxelsize;
  const bool on = fabsf(__fsub_rn(rx, __fmul_rn(fx, voxelsize))) <= tol &&
                  fabsf(__fsub_rn(ry, __fmul_rn(fy, voxelsize))) <= tol &&
                  fabsf(__fsub_rn(rz, __fmul_rn(fz, voxelsize))) <= tol && fx >= 0.f && fy >= 0.f && fz >= 0.f &&
                  fx < (float)kMaxCellsPerAxis && fy < (float)kMaxCellsPerAxis && fz < (float)kMaxCellsPerAxis;
  qx = on ? (int)fx : 0;
  qy = on ? (int)fy : 0;
  qz = on ? (int)fz : 0;
  return on;
}

__global__ void __launch_bounds__(kGridThreads)
grid_extent_kernel(unsigned char* ws, size_t per_set_bytes, const float* __restrict__ points, long long points_stride,
                   int n, float voxelsize, const int* __restrict__ filter, long long filter_stride) {
  __shared__ int s_dim[3];
  VoxelGridHeader* h = grid_header(ws, per_set_bytes, blockIdx.y);
  const float* pts = points + (long long)blockIdx.y * points_stride;
  const float gx = float_from_order_key(h->min_key[0]), gy = float_from_order_key(h->min_key[1]),
              gz = float_from_order_key(h->min_key[2]);
  if (threadIdx.x < 3) s_dim[threadIdx.x] = 0;
  __syncthreads();
  const int* keep = filter != nullptr ? filter + (long long)blockIdx.y * filter_stride : nullptr;
  int mx[3] = {0, 0, 0};
  bool bad = false;
  for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < n; i += gridDim.x * kGridThreads) {
    if (keep != nullptr && keep[i] < 0) continue;
    int q[3];
    bad = bad || !lattice_coords(pts, i, gx, gy, gz, voxelsize, q[0], q[1], q[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) mx[a] = max(mx[a], q[a] + 1);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int k = __reduce_max_sync(NSVF_FULL_MASK, mx[a]);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_dim[a], k);
  }
  if (__any_sync(NSVF_FULL_MASK, bad) && (threadIdx.x & 31) == 0) atomicOr(&h->bad, 1);
  __syncthreads();
  if (threadIdx.x < 3) atomicMax(&h->dims[threadIdx.x], s_dim[threadIdx.x]);
}

__global__ void __launch_bounds__(kGridThreads)
grid_clear_kernel(unsigned char* ws, size_t per_set_bytes) {
  VoxelGridHeader* h = grid_header(ws, per_set_bytes, blockIdx.y);
  const long long cells = (long long)h->dims[0] * h->dims[1] * h->dims[2];
  const bool fits = h->bad == 0 && cells > 0 && cells <= h->cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) h->ok = fits ? 1 : 0;
  if (!fits) return;
  int* cell = reinterpret_cast<int*>(h + 1);
  for (long long c = (long long)blockIdx.x * kGridThreads + threadIdx.x; c < cells;
       c += (long long)gridDim.x * kGridThreads)
    cell[c] = -1;
}

__global__ void __launch_bounds__(kGridThreads)
grid_scatter_kernel(unsigned char* ws, size_t per_set_bytes, const float* __restrict__ points, long long points_stride,
                    int n, float voxelsize, const int* __restrict__ filter, long long filter_stride) {
  VoxelGridHeader* h = grid_header(ws, per_set_bytes, blockIdx.y);
  if (h->ok == 0) return;
  const float* pts = points + (long long)blockIdx.y * points_stride;
  const float gx = float_from_order_key(h->min_key[0]), gy = float_from_order_key(h->min_key[1]),
              gz = float_from_order_key(h->min_key[2]);
  const int dy = h->dims[1], dz = h->dims[2];
  int* cell = reinterpret_cast<int*>(h + 1);
  const int* keep = filter != nullptr ? filter + (long long)blockIdx.y * filter_stride : nullptr;
  for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < n; i += gridDim.x * kGridThreads) {
    if (keep != nullptr && keep[i] < 0) continue;
    int qx, qy, qz;
    lattice_coords(pts, i, gx, gy, gz, voxelsize, qx, qy, qz);
    if (atomicCAS(&cell[((long long)qx * dy + qy) * dz + qz], -1, i) != -1) atomicOr(&h->bad, 1);   // two voxels, one cell
  }
}

int voxel_grid_build(cudaStream_t stream, int n_sets, int n, const float* points, long long points_stride,
                     float voxelsize, unsigned char* ws, size_t per_set_bytes, const int* filter, long long filter_stride) {
  const long long cap = (long long)(per_set_bytes - sizeof(VoxelGridHeader)) / 4;
  grid_init_kernel<<<(n_sets + 127) / 128, 128, 0, stream>>>(ws, per_set_bytes, n_sets, cap);
  NSVF_LAUNCH_OK("grid_init_kernel");
  int bx = (n + kGridThreads - 1) / kGridThreads;
  const int sms = num_sms();
  if (bx > 2 * sms) bx = 2 * sms;
  dim3 gp(bx, n_sets);
  grid_min_kernel<<<gp, kGridThreads, 0, stream>>>(ws, per_set_bytes, points, points_stride, n, filter, filter_stride);
  NSVF_LAUNCH_OK("grid_min_kernel");
  grid_extent_kernel<<<gp, kGridThreads, 0, stream>>>(ws, per_set_bytes, points, points_stride, n, voxelsize, filter,
                                                     filter_stride);
  NSVF_LAUNCH_OK("grid_extent_kernel");
  long long bc = (cap + kGridThreads * 8 - 1) / (kGridThreads * 8);
  if (bc > 4 * sms) bc = 4 * sms;
  grid_clear_kernel<<<dim3((unsigned)bc, n_sets), kGridThreads, 0, stream>>>(ws, per_set_bytes);
  NSVF_LAUNCH_OK("grid_clear_kernel");
  grid_scatter_kernel<<<gp, kGridThreads, 0, stream>>>(ws, per_set_bytes, points, points_stride, n, voxelsize, filter,
                                                      filter_stride);
  NSVF_LAUNCH_OK("grid_scatter_kernel");
  return 0;
}

// ---- the walk ------------------------------------------------------------------------------------------------------
enum { kWalkDepthSorted = 1, kWalkAnyHit = 2 };   // mode numbers of aabb_intersect.cu; index order (0) is not walked

__device__ __forceinline__ float pick3(int k, float a0, float a1, float a2) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }
__device__ __forceinline__ int pick3(int k, int a0, int a1, int a2) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }

struct WalkRay {
  float ox, oy, oz, ix, iy, iz;
  bool regular;
};

// the reference's test on voxel v's own box (centre -+ half voxel: the reference's first rounding)
__device__ __forceinline__ bool walk_test(const WalkRay& r, const float* __restrict__ pts, int v, float hv, float& tn,
                                          float& tf) {
  const float cx = __ldg(pts + (long long)v * 3), cy = __ldg(pts + (long long)v * 3 + 1),
              cz = __ldg(pts + (long long)v * 3 + 2);
  const float lx = __fsub_rn(cx, hv), ly = __fsub_rn(cy, hv), lz = __fsub_rn(cz, hv);
  const float hx = __fadd_rn(cx, hv), hy = __fadd_rn(cy, hv), hz = __fadd_rn(cz, hv);
  if (r.regular) {   // no NaN possible: fmin/fmax ordering == the reference's swap
    const float a0 = __fmul_rn(__fsub_rn(lx, r.ox), r.ix), b0 = __fmul_rn(__fsub_rn(hx, r.ox), r.ix);
    const float a1 = __fmul_rn(__fsub_rn(ly, r.oy), r.iy), b1 = __fmul_rn(__fsub_rn(hy, r.oy), r.iy);
    const float a2 = __fmul_rn(__fsub_rn(lz, r.oz), r.iz), b2 = __fmul_rn(__fsub_rn(hz, r.oz), r.iz);
    tn = fmaxf(fmaxf(0.0f, fminf(a0, b0)), fmaxf(fminf(a1, b1), fminf(a2, b2)));
    tf = fminf(fminf(100000.0f, fmaxf(a0, b0)), fminf(fmaxf(a1, b1), fmaxf(a2, b2)));
    return tn <= tf;
  }
  return slab_exact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, lx, ly, lz, hx, hy, hz, tn, tf);
}

// Per-ray hit list under construction, living in the ray's own output row (global memory).
template <int MODE>
struct WalkRow {
  int* idx;
  float* dmin;
  float* dmax;
  const int* prio;   // nullptr: ties and truncation go by voxel index (aabb); else by prio[v] (octree: DFS rank of the leaf)
  int n_max, cnt;
  bool unsorted;
  float last_tn;
  int last_v;        // key of the last hit

  __device__ __forceinline__ int key(int v) const { return prio != nullptr ? __ldg(prio + v) : v; }
  __device__ __forceinline__ void reset(int n_max_) {
    n_max = n_max_;
    cnt = 0;
    unsorted = false;
    last_tn = 0.f;
    last_v = -1;
  }
  __device__ __forceinline__ void note_order(int v, float tn) {
    const int k = key(v);
    if (cnt > 0 && (tn < last_tn || (tn == last_tn && k < last_v))) unsorted = true;
    last_tn = tn;
    last_v = k;
  }
  __device__ __forceinline__ void add(int v, float tn, float tf) {
    if (MODE == kWalkAnyHit) { cnt = 1; return; }
    if (cnt < n_max) {
      idx[cnt] = v;
      dmin[cnt] = tn;
      dmax[cnt] = tf;
      note_order(v, tn);
      ++cnt;
      return;
    }
    // full: the reference keeps the n_max FIRST hits of its own visiting order (ascending voxel index for the linear
    // scan, DFS order for the octree): the n_max smallest keys
    int worst = -1, at = 0;
    for (int s = 0; s < n_max; ++s) {
      const int w = key(idx[s]);
      if (w > worst) { worst = w; at = s; }
    }
    if (key(v) < worst) {
      idx[at] = v; dmin[at] = tn; dmax[at] = tf;
      unsorted = true;
    }
  }
  // insertion sort of the row by (depth, key): only after an order violation was seen
  __device__ __forceinline__ void repair() {
    for (int s = 1; s < cnt; ++s) {
      const int v = idx[s];
      const int kv = key(v);
      const float tn = dmin[s], tf = dmax[s];
      int t = s - 1;
      while (t >= 0) {
        const float e = dmin[t];
        const int w = idx[t];
        if (!(e > tn || (e == tn && key(w) > kv))) break;
        idx[t + 1] = w; dmin[t + 1] = e; dmax[t + 1] = dmax[t];
        --t;
      }
      idx[t + 1] = v; dmin[t + 1] = tn; dmax[t + 1] = tf;
    }
  }
};

// The ray in lattice coordinates (cell i covers [i, i + 1)), axes permuted so that m is the dominant direction.
struct WalkPath {
  float um0, up0, uq0, slope_p, slope_q, eps;
  int nm, np, nq, sm, sp, sq;     // extents and cell strides of the permuted axes
  int i, i_end, step;             // layers i, i + step, ... up to and including i_end
  bool fwd, p_up, q_up;
  bool trusted;                   // false: scan all voxels instead (see file header)

  __device__ __forceinline__ bool more() const { return fwd ? i <= i_end : i >= i_end; }

  __device__ __forceinline__ void setup(const VoxelGridHeader* __restrict__ h, float voxelsize, const WalkRay& r,
                                        float dx, float dy, float dz) {
    const int nx = h->dims[0], ny = h->dims[1], nz = h->dims[2];
    const float gx = float_from_order_key(h->min_key[0]), gy = float_from_order_key(h->min_key[1]),
                gz = float_from_order_key(h->min_key[2]);
    const float ux = __fdiv_rn(r.ox - gx, voxelsize) + 0.5f, uy = __fdiv_rn(r.oy - gy, voxelsize) + 0.5f,
                uz = __fdiv_rn(r.oz - gz, voxelsize) + 0.5f;
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const float am = fmaxf(ax, fmaxf(ay, az));
    const bool finite = fabsf(r.ox) <= 3.0e38f && fabsf(r.oy) <= 3.0e38f && fabsf(r.oz) <= 3.0e38f && ax <= 3.0e38f &&
                        ay <= 3.0e38f && az <= 3.0e38f;
    const float far = fmaxf(fabsf(ux), fmaxf(fabsf(uy), fabsf(uz)));
    trusted = finite && am >= 1.0e-18f && am <= 1.0e18f && far <= 1.0e5f;
    i = 0; i_end = -1; step = 1; fwd = true;     // no layers
    if (!trusted) return;
    const int m = (ay > ax) ? ((az > ay) ? 2 : 1) : ((az > ax) ? 2 : 0);
    const int p = m == 2 ? 0 : m + 1, q = p == 2 ? 0 : p + 1;
    um0 = pick3(m, ux, uy, uz); up0 = pick3(p, ux, uy, uz); uq0 = pick3(q, ux, uy, uz);
    const float dm = pick3(m, dx, dy, dz), dp = pick3(p, dx, dy, dz), dq = pick3(q, dx, dy, dz);
    nm = pick3(m, nx, ny, nz); np = pick3(p, nx, ny, nz); nq = pick3(q, nx, ny, nz);
    const int sx = ny * nz, sy = nz;
    sm = pick3(m, sx, sy, 1); sp = pick3(p, sx, sy, 1); sq = pick3(q, sx, sy, 1);
    const float inv_dm = __fdiv_rn(1.0f, dm);
    slope_p = dp * inv_dm; slope_q = dq * inv_dm;   // |slope| <= 1
    // candidate margin in cells: lattice tolerance + rounding of the reference's test and of this walk (both grow with
    // the distance of the origin, ~1e-7 relative)
    eps = 4.0e-3f + 4.0e-6f * (fabsf(um0) + fabsf(up0) + fabsf(uq0) + (float)(nm + np + nq));
    p_up = dp >= 0.0f; q_up = dq >= 0.0f;
    // layers in which both minor coordinates can be inside the lattice (two cells of slack; an almost constant minor
    // coordinate drifts by less than two cells over the representable range)
    float lo_u = -1.0f, hi_u = (float)nm + 1.0f;
    const float s2[2] = {slope_p, slope_q}, u2[2] = {up0, uq0}, n2[2] = {(float)np, (float)nq};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (fabsf(s2[k]) >= 1.0e-6f) {
        const float t0 = um0 + __fdiv_rn(-2.0f - u2[k], s2[k]), t1 = um0 + __fdiv_rn(n2[k] + 2.0f - u2[k], s2[k]);
        lo_u = fmaxf(lo_u, fminf(t0, t1));
        hi_u = fminf(hi_u, fmaxf(t0, t1));
      } else if (u2[k] < -4.0f || u2[k] > n2[k] + 4.0f) {
        hi_u = -2.0f;
      }
    }
    if (!(lo_u <= hi_u)) return;
    const int i_lo = max(0, (int)floorf(lo_u) - 1), i_hi = min(nm - 1, (int)floorf(hi_u) + 1);
    fwd = dm > 0.0f;
    step = fwd ? 1 : -1;
    i = fwd ? max(i_lo, (int)floorf(um0 - eps)) : min(i_hi, (int)floorf(um0 + eps));
    i_end = fwd ? i_hi : i_lo;
  }

  // the occupied candidate cells of layer i in travel order (so that hits come out sorted by entry depth in all but
  // degenerate cases); visit(v) returns true to stop
  template <class F>
  __device__ __forceinline__ bool layer(const int* __restrict__ cell, F&& visit) const {
    // the part of the ray (t >= 0) inside this layer, in the dominant coordinate
    float ua = (float)i - eps, ub = (float)(i + 1) + eps;
    if (fwd) ua = fmaxf(ua, um0 - eps);
    else ub = fminf(ub, um0 + eps);
    const float ra = ua - um0, rb = ub - um0;
    const float pa = fmaf(slope_p, ra, up0), pb = fmaf(slope_p, rb, up0);
    const int jp0 = max(0, (int)floorf(fminf(pa, pb) - eps)), jp1 = min(np - 1, (int)floorf(fmaxf(pa, pb) + eps));
    if (jp0 > jp1) return false;
    const float qa = fmaf(slope_q, ra, uq0), qb = fmaf(slope_q, rb, uq0);
    const int jq0 = max(0, (int)floorf(fminf(qa, qb) - eps)), jq1 = min(nq - 1, (int)floorf(fmaxf(qa, qb) + eps));
    if (jq0 > jq1) return false;
    const int base = i * sm;
    for (int a = 0; a <= jp1 - jp0; ++a) {
      const int jp = p_up ? jp0 + a : jp1 - a;
      for (int b = 0; b <= jq1 - jq0; ++b) {
        const int jq = q_up ? jq0 + b : jq1 - b;
        const int v = __ldg(cell + base + jp * sp + jq * sq);
        if (v >= 0 && visit(v)) return true;
      }
    }
    return false;
  }
};

// One ray start to finish by one thread, hits written straight into its global row: the any-hit query, and the sorted
// query's slow path (rows that overflow n_max, rays the walk does not trust).
template <int MODE>
__device__ __forceinline__ void walk_direct(const VoxelGridHeader* __restrict__ h, const float* __restrict__ pts, int n,
                                            float hv, const WalkRay& r, WalkPath path, WalkRow<MODE>& row) {
  const int* __restrict__ cell = reinterpret_cast<const int*>(h + 1);
  auto visit = [&](int v) -> bool {
    float tn, tf;
    if (!walk_test(r, pts, v, hv, tn, tf)) return false;
    row.add(v, tn, tf);
    return MODE == kWalkAnyHit;
  };
  if (!path.trusted) {   // scan all voxels in index order like the reference
    for (int v = 0; v < n; ++v)
      if (visit(v)) return;
    return;
  }
  for (; path.more(); path.i += path.step)
    if (path.layer(cell, visit)) return;
}

constexpr int kRing = 16, kRingLd = 17;   // staged hits per ray, padded row stride (bank spread)

template <int MODE>
__global__ void __launch_bounds__(kWalkThreads)
grid_walk_kernel(const unsigned char* __restrict__ ws, size_t per_set_bytes, const float* __restrict__ points,
                 long long points_stride, int n, float voxelsize, long long rays_per_set, int n_max, float empty_depth,
                 const float* __restrict__ ray_start, const float* __restrict__ ray_dir, int* __restrict__ out_idx,
                 float* __restrict__ out_min, float* __restrict__ out_max, unsigned char* __restrict__ out_hit,
                 WalkOctree oct) {
  // the hierarchy / octree kernels take this voxel set when it is not a lattice (or the octree is not a proper tree)
  const bool mine = voxel_grid_usable(ws, per_set_bytes, blockIdx.y) && (oct.veto == nullptr || oct.veto[blockIdx.y] == 0);
  if (oct.active != nullptr && blockIdx.x == 0 && threadIdx.x == 0) oct.active[blockIdx.y] = mine ? 1 : 0;
  if (!mine) return;
  const int* prio = oct.rank != nullptr ? oct.rank + (long long)blockIdx.y * oct.rank_stride : nullptr;
  constexpr int kWarps = kWalkThreads / 32;
  constexpr int kStage = MODE == kWalkDepthSorted ? kWarps * 32 * kRingLd : 1;
  __shared__ int s_idx[kStage];
  __shared__ float s_min[kStage], s_max[kStage];
  const VoxelGridHeader* h = reinterpret_cast<const VoxelGridHeader*>(ws + (size_t)blockIdx.y * per_set_bytes);
  const int* __restrict__ cell = reinterpret_cast<const int*>(h + 1);
  const float* pts = points + (long long)blockIdx.y * points_stride;
  const float hv = voxelsize * 0.5f;   // reference: float half_voxel = voxelsize * 0.5 (exact)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray_base = (long long)blockIdx.y * rays_per_set;
  const long long n_tiles = (rays_per_set + 31) / 32;
  for (long long tile = (long long)blockIdx.x * kWarps + warp; tile < n_tiles; tile += (long long)gridDim.x * kWarps) {
    const long long rr = tile * 32 + lane;
    const bool live = rr < rays_per_set;
    const long long ray = ray_base + (live ? rr : rays_per_set - 1);
    WalkRay r;
    r.ox = ray_start[ray * 3 + 0]; r.oy = ray_start[ray * 3 + 1]; r.oz = ray_start[ray * 3 + 2];
    const float dx = ray_dir[ray * 3 + 0], dy = ray_dir[ray * 3 + 1], dz = ray_dir[ray * 3 + 2];
    r.ix = ref_rcp(dx); r.iy = ref_rcp(dy); r.iz = ref_rcp(dz);
    r.regular = regular_component(r.ox, r.ix) && regular_component(r.oy, r.iy) && regular_component(r.oz, r.iz);
    WalkPath path;
    path.setup(h, voxelsize, r, dx, dy, dz);
    WalkRow<MODE> row;
    row.reset(n_max);
    row.prio = prio;
    // octree queries: the reference's answer for rays with NaN paths (and rays this walk does not trust) depends on the
    // loose internal boxes, so those rays are left to the traversal kernel (defer[ray] = 1, nothing written here)
    const bool deferred = oct.defer != nullptr && live && !(r.regular && path.trusted);
    if (oct.defer != nullptr && live) oct.defer[ray] = deferred ? 1 : 0;
    row.idx = MODE == kWalkAnyHit ? nullptr : out_idx + ray * n_max;
    row.dmin = MODE == kWalkAnyHit ? nullptr : out_min + ray * n_max;
    row.dmax = MODE == kWalkAnyHit ? nullptr : out_max + ray * n_max;

    if constexpr (MODE == kWalkAnyHit) {
      if (live) {
        walk_direct<MODE>(h, pts, n, hv, r, path, row);
        out_hit[ray] = row.cnt > 0;
      }
    } else {
      // The 32 rays of the tile advance one layer at a time, in step.  Hits are staged in a small ring per ray in
      // shared memory; whenever some ray has 8 waiting, the warp writes every ray's staged hits to its row with 8
      // lanes per row (32 contiguous bytes) — thread-per-ray stores would touch 32 different sectors per instruction.
      int* ring_i = s_idx + (warp * 32 + lane) * kRingLd;
      float* ring_a = s_min + (warp * 32 + lane) * kRingLd;
      float* ring_b = s_max + (warp * 32 + lane) * kRingLd;
      const long long tile_row0 = (ray_base + tile * 32) * n_max;
      int flushed = 0;
      bool slow = live && !path.trusted && !deferred;      // redo on the slow path: untrusted ray, row or ring overflow
      bool walking = live && path.trusted && !deferred && path.more();
      auto flush = [&](int at_least) {        // warp-uniform; writes up to 8 staged hits of every ray
        while (__any_sync(NSVF_FULL_MASK, row.cnt - flushed >= at_least)) {
#pragma unroll 1
          for (int k0 = 0; k0 < 32; k0 += 4) {
            const int k = k0 + (lane >> 3), t = lane & 7;
            const int f = __shfl_sync(NSVF_FULL_MASK, flushed, k), c = __shfl_sync(NSVF_FULL_MASK, row.cnt, k);
            if (t < c - f) {
              const int src = (warp * 32 + k) * kRingLd + ((f + t) & (kRing - 1));
              const long long dst = tile_row0 + (long long)k * n_max + f + t;
              out_idx[dst] = s_idx[src];
              out_min[dst] = s_min[src];
              out_max[dst] = s_max[src];
            }
          }
          flushed += min(row.cnt - flushed, 8);
          __syncwarp();
        }
      };
      while (__any_sync(NSVF_FULL_MASK, walking)) {
        if (walking) {
          path.layer(cell, [&](int v) -> bool {
            float tn, tf;
            if (!walk_test(r, pts, v, hv, tn, tf)) return false;
            if (row.cnt >= n_max || row.cnt - flushed >= kRing) { slow = true; return true; }
            const int s = row.cnt & (kRing - 1);
            ring_i[s] = v; ring_a[s] = tn; ring_b[s] = tf;
            row.note_order(v, tn);
            ++row.cnt;
            return false;
          });
          path.i += path.step;
          walking = !slow && path.more();
        }
        __syncwarp();
        flush(8);
      }
      flush(1);
      if (slow) {          // rare: start this ray over, thread-per-ray into the global row
        row.reset(n_max);
        path.setup(h, voxelsize, r, dx, dy, dz);
        walk_direct<MODE>(h, pts, n, hv, r, path, row);
      }
      if (live && row.unsorted) row.repair();
      if (live && !deferred && out_hit != nullptr) out_hit[ray] = row.cnt > 0;
      // tails of the 32 rows of this tile: coalesced -1 / fill depth
      const int my_cnt = (live && !deferred) ? row.cnt : n_max;
      for (int k = 0; k < 32; ++k) {
        const int c = __shfl_sync(NSVF_FULL_MASK, my_cnt, k);
        const long long r0 = tile_row0 + (long long)k * n_max;
        for (int s = c + lane; s < n_max; s += 32) {
          out_idx[r0 + s] = -1;
          out_min[r0 + s] = empty_depth;
          out_max[r0 + s] = empty_depth;
        }
      }
      __syncwarp();
    }
  }
}

int voxel_grid_walk(cudaStream_t stream, int mode, const unsigned char* ws, size_t per_set_bytes, int n_sets, int n,
                    const float* points, long long points_stride, float voxelsize, long long rays_per_set, int n_max,
                    float empty_depth, const float* ray_start, const float* ray_dir, int* idx, float* min_depth,
                    float* max_depth, unsigned char* hit, const WalkOctree* octree) {
  WalkOctree oct{};
  if (octree != nullptr) oct = *octree;
  const long long tiles = (rays_per_set + 31) / 32;
  long long want = (tiles + kWalkThreads / 32 - 1) / (kWalkThreads / 32), cap = (long long)num_sms() * 16;
  if (n_sets > 1) cap = (cap + n_sets - 1) / n_sets;
  dim3 grid((unsigned)(want < cap ? want : cap), n_sets);
  if (grid.x < 1) grid.x = 1;
#define NSVF_WALK(MODE, NAME)                                                                                       \
  NSVF_TIMED_LAUNCH(NAME, stream,                                                                                   \
                    (grid_walk_kernel<MODE><<<grid, kWalkThreads, 0, stream>>>(                                     \
                        ws, per_set_bytes, points, points_stride, n, voxelsize, rays_per_set, n_max, empty_depth,   \
                        ray_start, ray_dir, idx, min_depth, max_depth, hit, oct)))
  NSVF_REQUIRE(mode == kWalkDepthSorted || mode == kWalkAnyHit, "voxel_grid_walk: mode must be 1 (sorted) or 2 (any hit)");
  if (mode == kWalkDepthSorted && octree != nullptr) NSVF_WALK(kWalkDepthSorted, "svo_intersect_sorted_kernel");
  else if (mode == kWalkDepthSorted) NSVF_WALK(kWalkDepthSorted, "aabb_intersect_sorted_kernel");
  else NSVF_WALK(kWalkAnyHit, "aabb_hit_mask_kernel");
#undef NSVF_WALK
  return 0;
}

}  // namespace nsvf
