/* nsvf_oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build,
 * load or call this file; the product (nsvf_b200/) never does and has no CPU path.
 *
 * Every function restates one reference function, loop for loop, and cites it.  Paths are relative
 * to the reference checkout (facebookresearch/NSVF).  Build: gcc -O2 -ffp-contract=off -fopenmp
 * (contraction is off so that the only fused multiply-adds are the explicit fmaf() calls placed
 * where nvcc's default -fmad=true contracts the reference expression).
 *
 * Parity pinning: the reference has no tests or golden vectors.  This oracle is
 * pinned against (1) the reference's own CUDA kernels (oracle/_ref/ref_ext.so, the unmodified
 * fairnr/clib compiled for sm_100a) executed on the GPU box — fixtures under tests/golden/ made by
 * tests/golden/make_gpu_golden.py — and (2) the reference's Python modules imported unmodified in
 * the build container — fixtures made by tests/golden/make_cpu_golden.py.
 *
 * One deliberate difference from the device code: the reference computes 1/dir with
 * __fdividef (MUFU.RCP based, <= 1 ulp from IEEE).  The oracle takes the reciprocals as an optional
 * input (`inv_dir`, produced on the GPU with the same intrinsic) and otherwise uses IEEE 1.0f/d;
 * with IEEE reciprocals, rays that graze a voxel face within 1 ulp may differ.
 *
 * Where the reference reads or writes out of bounds the oracle
 * defines the behaviour exactly like the product does: a read past the end of a tensor yields -1,
 * a sample that would land at s >= max_steps is dropped.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ---- RayAABBIntersection, fairnr/clib/src/intersect_gpu.cu:73-122 -------------------------------- */
static int ray_aabb(const float ori[3], const float inv[3], const float center[3], float half_voxel,
                    float* t_near, float* t_far) {
  float f_low = 0.0f, f_high = 100000.0f;
  for (int d = 0; d < 3; ++d) {
    float inv_ray_dir = inv[d], start = ori[d], aabb = center[d];
    float f_dim_low = (aabb - half_voxel - start) * inv_ray_dir;
    float f_dim_high = (aabb + half_voxel - start) * inv_ray_dir;
    if (f_dim_high < f_dim_low) { float t = f_dim_low; f_dim_low = f_dim_high; f_dim_high = t; }
    if (f_dim_high < f_low) return 0;
    if (f_dim_low > f_high) return 0;
    f_low = (f_dim_low > f_low) ? f_dim_low : f_low;
    f_high = (f_dim_high < f_high) ? f_dim_high : f_high;
    if (f_low > f_high) return 0;
  }
  *t_near = f_low;
  *t_far = f_high;
  return f_low > -1.0f; /* reference: if (depths.x > -1.0f) */
}

static void ray_inv(const float* ray_dir, const float* inv_dir, long long j, float inv[3]) {
  for (int d = 0; d < 3; ++d) inv[d] = inv_dir ? inv_dir[j * 3 + d] : 1.0f / ray_dir[j * 3 + d];
}

/* aabb_intersect_point_kernel, intersect_gpu.cu:125-167 + output init of intersect.cpp:61-69 */
ORACLE_API void oracle_aabb_intersect(int b, int n, int m, float voxelsize, int n_max, const float* ray_start,
                                      const float* ray_dir, const float* inv_dir, const float* points,
                                      long long points_batch_stride, int* idx, float* min_depth,
                                      float* max_depth) {
  const float half_voxel = voxelsize * 0.5f;
#pragma omp parallel for schedule(dynamic, 64)
  for (long long r = 0; r < (long long)b * m; ++r) {
    const float* pts = points + (r / m) * points_batch_stride;
    float inv[3];
    ray_inv(ray_dir, inv_dir, r, inv);
    for (int l = 0; l < n_max; ++l) { idx[r * n_max + l] = -1; min_depth[r * n_max + l] = 0.f; max_depth[r * n_max + l] = 0.f; }
    for (int k = 0, cnt = 0; k < n && cnt < n_max; ++k) {
      float tn, tf;
      if (ray_aabb(ray_start + r * 3, inv, pts + (long long)k * 3, half_voxel, &tn, &tf)) {
        idx[r * n_max + cnt] = k;
        min_depth[r * n_max + cnt] = tn;
        max_depth[r * n_max + cnt] = tf;
        ++cnt;
      }
    }
  }
}

/* svo_intersect_point_kernel, intersect_gpu.cu:170-237. returns the number of rays that hit the
 * reference's assert((ptr < 256)) (they are stopped instead). */
ORACLE_API int oracle_svo_intersect(int b, int T, int m, float voxelsize, int n_max, const float* ray_start,
                                    const float* ray_dir, const float* inv_dir, const float* points,
                                    const int* children, long long tree_batch_stride_nodes, int* idx,
                                    float* min_depth, float* max_depth) {
  const float half_voxel = voxelsize * 0.5f;
  int overflow = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : overflow)
  for (long long r = 0; r < (long long)b * m; ++r) {
    const float* pts = points + (r / m) * tree_batch_stride_nodes * 3;
    const int* ch = children + (r / m) * tree_batch_stride_nodes * 9;
    float inv[3];
    ray_inv(ray_dir, inv_dir, r, inv);
    for (int l = 0; l < n_max; ++l) { idx[r * n_max + l] = -1; min_depth[r * n_max + l] = 0.f; max_depth[r * n_max + l] = 0.f; }
    int stack[256];
    int ptr = 0, cnt = 0;
    stack[0] = T - 1;
    while (ptr > -1 && cnt < n_max) {
      int k = stack[ptr];
      float tn, tf;
      int hit = ray_aabb(ray_start + r * 3, inv, pts + (long long)k * 3, half_voxel * (float)ch[(long long)k * 9 + 8], &tn, &tf);
      ptr--;
      if (hit) {
        if (ch[(long long)k * 9 + 8] == 1) {
          idx[r * n_max + cnt] = k;
          min_depth[r * n_max + cnt] = tn;
          max_depth[r * n_max + cnt] = tf;
          ++cnt;
          continue;
        }
        if (ptr + 8 >= 256) { overflow += 1; break; }
        for (int u = 0; u < 8; u++)
          if (ch[(long long)k * 9 + u] > -1) { ptr++; stack[ptr] = ch[(long long)k * 9 + u]; }
      }
    }
  }
  return overflow;
}

/* uniform_ray_sampling_kernel, fairnr/clib/src/sample_gpu.cu:15-106 + output init of sample.cpp:40-48 */
ORACLE_API void oracle_uniform_ray_sampling(int b, int num_rays, int max_hits, int max_steps, float step_size,
                                            const int* pts_idx, const float* min_depth, const float* max_depth,
                                            const float* uniform_noise, int* sampled_idx, float* sampled_depth,
                                            float* sampled_dists) {
  const long long total_rays = (long long)b * num_rays, total_hits = total_rays * max_hits;
#pragma omp parallel for schedule(dynamic, 64)
  for (long long j = 0; j < total_rays; ++j) {
    const long long H = j * max_hits, K = j * max_steps;
    for (int t = 0; t < max_steps; ++t) { sampled_idx[K + t] = -1; sampled_depth[K + t] = 0.f; sampled_dists[K + t] = 0.f; }
    int s = 0, ucur = 0, umin = 0, umax = 0;
    float last_min_depth = 0.f, last_max_depth = 0.f, curr_depth = 0.f;
    while (1) {
      if ((umax == max_hits) || (ucur == max_steps) || (pts_idx[H + umax] == -1)) break;
      last_min_depth = (umin < max_hits) ? min_depth[H + umin] : 10000.0f;
      last_max_depth = (umax < max_hits) ? max_depth[H + umax] : 10000.0f;
      if (ucur < max_steps) /* nvcc contracts min_depth[H] + (float(ucur) + noise) * step_size into one FFMA */
        curr_depth = fmaf((float)ucur + uniform_noise[K + ucur], step_size, min_depth[H]);
      if (s >= max_steps) break; /* reference writes past the row */
      if ((last_max_depth <= curr_depth) && (last_max_depth <= last_min_depth)) {
        sampled_depth[K + s] = last_max_depth;
        sampled_idx[K + s] = pts_idx[H + umax];
        umax++; s++; continue;
      }
      if ((curr_depth <= last_min_depth) && (curr_depth <= last_max_depth)) {
        long long f = H + umin - 1;
        sampled_depth[K + s] = curr_depth;
        sampled_idx[K + s] = (f >= 0 && f < total_hits) ? pts_idx[f] : -1;
        ucur++; s++; continue;
      }
      if ((last_min_depth <= curr_depth) && (last_min_depth <= last_max_depth)) {
        sampled_depth[K + s] = last_min_depth;
        sampled_idx[K + s] = (umin < max_hits) ? pts_idx[H + umin] : -1;
        umin++; s++; continue;
      }
      break; /* NaN inputs: the reference spins forever */
    }
    int step = 0;
    umin = 0; umax = 0;
    for (ucur = 0; ucur < max_steps - 1; ucur++) {
      if (sampled_idx[K + ucur + 1] == -1) break;
      float l_depth = sampled_depth[K + ucur], r_depth = sampled_depth[K + ucur + 1];
      sampled_depth[K + ucur] = (l_depth + r_depth) * .5f;
      sampled_dists[K + ucur] = (r_depth - l_depth);
      if ((umin < max_hits) && (sampled_depth[K + ucur] >= min_depth[H + umin]) && (pts_idx[H + umin] > -1)) umin++;
      if ((umax < max_hits) && (sampled_depth[K + ucur] >= max_depth[H + umax]) && (pts_idx[H + umax] > -1)) umax++;
      if ((umax == max_hits) || (pts_idx[H + umax] == -1)) break;
      if ((umin - 1 == umax) && (sampled_dists[K + ucur] > 0)) {
        sampled_depth[K + step] = sampled_depth[K + ucur];
        sampled_dists[K + step] = sampled_dists[K + ucur];
        sampled_idx[K + step] = sampled_idx[K + ucur];
        step++;
      }
    }
    for (int t = step; t < max_steps; t++) sampled_idx[K + t] = -1;
  }
}

/* inverse_cdf_sampling_kernel, sample_gpu.cu:108-202 + output init of sample.cpp:80-88.
 * The tensors are one contiguous [b, num_rays, *] call of the reference kernel. */
ORACLE_API void oracle_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps, float fixed_step_size,
                                            const int* pts_idx, const float* min_depth, const float* max_depth,
                                            const float* uniform_noise, const float* probs, const float* steps,
                                            int* sampled_idx, float* sampled_depth, float* sampled_dists) {
  const long long total_rays = (long long)b * num_rays, total_hits = total_rays * max_hits;
#define EMIT(I, DIST, DEPTH)                                                      \
  do {                                                                            \
    if (s < max_steps) { sampled_idx[K + s] = (I); sampled_dists[K + s] = (DIST); sampled_depth[K + s] = (DEPTH); } \
  } while (0)
#pragma omp parallel for schedule(dynamic, 64)
  for (long long j = 0; j < total_rays; ++j) {
    const long long batch = j / num_rays;
    const int* row0 = pts_idx + batch * num_rays * max_hits; /* `pts_idx[curr_bin]` of line 194 */
    const long long H = j * max_hits, K = j * max_steps;
    for (int t = 0; t < max_steps; ++t) { sampled_idx[K + t] = -1; sampled_depth[K + t] = 0.f; sampled_dists[K + t] = 0.f; }
    int curr_bin = 0, s = 0;
    float curr_min_depth = min_depth[H], curr_max_depth = max_depth[H];
    float curr_min_cdf = 0, curr_max_cdf = probs[H];
    float step_size = (float)(1.0 / (double)steps[j]);
    float z_low = curr_min_depth;
    int total_steps = (int)ceilf(steps[j]);
    int done = 0;
    if (fixed_step_size > 0.0f) step_size = fixed_step_size;
    for (int curr_step = 0; curr_step < total_steps; curr_step++) {
      int ns = curr_step < max_steps ? curr_step : max_steps - 1; /* reference reads past the noise row */
      float curr_cdf = ((float)curr_step + uniform_noise[K + ns]) * step_size;
      while (curr_cdf > curr_max_cdf) {
        EMIT(pts_idx[H + curr_bin], curr_max_depth - z_low, (curr_max_depth + z_low) * .5f);
        curr_bin++; s++;
        if ((curr_bin >= max_hits) || (pts_idx[H + curr_bin] == -1)) { done = 1; break; }
        curr_min_depth = min_depth[H + curr_bin];
        curr_max_depth = max_depth[H + curr_bin];
        curr_min_cdf = curr_max_cdf;
        curr_max_cdf = curr_max_cdf + probs[H + curr_bin];
        z_low = curr_min_depth;
      }
      if (done) break;
      float u = (curr_cdf - curr_min_cdf) / (curr_max_cdf - curr_min_cdf);
      float z = fmaf(u, curr_max_depth - curr_min_depth, curr_min_depth); /* contracted by nvcc */
      EMIT(pts_idx[H + curr_bin], z - z_low, (z + z_low) * .5f);
      z_low = z; s++;
    }
    /* `(~done)` is always true (bitwise not of a bool) */
    while (z_low < curr_max_depth) {
      long long f = H + curr_bin;
      EMIT(f < total_hits ? pts_idx[f] : -1, curr_max_depth - z_low, (curr_max_depth + z_low) * .5f);
      curr_bin++; s++;
      if ((curr_bin >= max_hits) || (row0[curr_bin] == -1)) break;
      curr_min_depth = min_depth[H + curr_bin];
      curr_max_depth = max_depth[H + curr_bin];
      z_low = curr_min_depth;
    }
  }
#undef EMIT
}

/* ---- build_octree, fairnr/clib/src/octree.cpp:13-135 ------------------------------------------------ */
typedef struct ONode {
  float c[3];
  int depth, index;
  struct ONode* ch[8];
} ONode;

static ONode* onode_new(const float* c, int d, int i) {
  ONode* p = (ONode*)calloc(1, sizeof(ONode));
  p->c[0] = c[0]; p->c[1] = c[1]; p->c[2] = c[2];
  p->depth = d; p->index = i;
  return p;
}
static void onode_insert(ONode* p, const float* pt, int index) { /* EasyOctree::insert :62-77 */
  int bit[3];
  for (int a = 0; a < 3; ++a) bit[a] = pt[a] > p->c[a];
  int idx = bit[0] + 2 * bit[1] + 4 * bit[2];
  if (p->depth == 0) {
    p->ch[idx] = onode_new(pt, -1, index); /* the reference leaks the previous leaf, if any */
  } else {
    if (p->ch[idx] == NULL) {
      int length = 1 << (p->depth - 1);
      float nc[3];
      for (int a = 0; a < 3; ++a) nc[a] = p->c[a] + (float)((2 * bit[a] - 1) * length);
      p->ch[idx] = onode_new(nc, p->depth - 1, -1);
    }
    onode_insert(p->ch[idx], pt, index);
  }
}
static void onode_count(ONode* p, long long* total, long long* terminal) { /* :79-91 */
  for (int i = 0; i < 8; i++) if (p->ch[i]) onode_count(p->ch[i], total, terminal);
  *total += 1;
  if (p->depth == -1) *terminal += 1;
}
static void onode_free(ONode* p) {
  for (int i = 0; i < 8; i++) if (p->ch[i]) onode_free(p->ch[i]);
  free(p);
}
static ONode* g_root = NULL;
static long long g_total = 0;

ORACLE_API long long oracle_octree_build(const float* center, const long long* points, long long n, int depth) {
  if (g_root) onode_free(g_root);
  g_root = onode_new(center, depth, -1);
  for (long long k = 0; k < n; ++k) {
    float pt[3] = {(float)points[k * 3], (float)points[k * 3 + 1], (float)points[k * 3 + 2]};
    onode_insert(g_root, pt, (int)k);
  }
  long long total = 0, terminal = 0;
  onode_count(g_root, &total, &terminal);
  g_total = total;
  return total;
}
ORACLE_API void oracle_octree_flatten(int* centers, int* children) { /* EasyOctree::finalize :93-123 */
  long long T = g_total;
  memset(centers, 0, sizeof(int) * T * 3);
  for (long long i = 0; i < T * 9; ++i) children[i] = -1;
  ONode** queue = (ONode**)malloc(sizeof(ONode*) * (size_t)(T + 8));
  long long head = 0, tail = 0, node_idx = T - 1;
  g_root->index = (int)node_idx;
  queue[tail++] = g_root;
  while (head < tail) {
    ONode* nd = queue[head++];
    for (int i = 0; i < 8; i++) {
      if (nd->ch[i]) {
        if (nd->ch[i]->depth > -1) { node_idx--; nd->ch[i]->index = (int)node_idx; }
        queue[tail++] = nd->ch[i];
        if (nd->index >= 0 && nd->index < T) children[(long long)nd->index * 9 + i] = nd->ch[i]->index;
      }
    }
    if (nd->index >= 0 && nd->index < T) {
      children[(long long)nd->index * 9 + 8] = 1 << (nd->depth + 1);
      for (int a = 0; a < 3; ++a) centers[(long long)nd->index * 3 + a] = (int)nd->c[a];
    }
  }
  free(queue);
  onode_free(g_root);
  g_root = NULL;
}

/* ---- trilinear interpolation: SparseVoxelEncoder.forward, fairnr/modules/encoder.py:582-590 with
 * offset_points (geometry.py:229-238) and trilinear_interp (geometry.py:195-200), float32 ---------- */
static void tri_weights(const float* xyz, const float* c, float voxel_size, float w[8], float p[3]) {
  for (int a = 0; a < 3; ++a) p[a] = (xyz[a] - c[a]) / voxel_size + .5f;
  for (int j = 0; j < 8; ++j) { /* q = offset order: x slowest, z fastest; w = prod_a (p*q + (1-p)*(1-q)) */
    float q[3] = {(float)((j >> 2) & 1), (float)((j >> 1) & 1), (float)(j & 1)};
    float t[3];
    for (int a = 0; a < 3; ++a) t[a] = p[a] * q[a] + (1.f - p[a]) * (1.f - q[a]);
    w[j] = (t[0] * t[1]) * t[2];
  }
}
ORACLE_API void oracle_trilinear_fwd(long long M, int D, const int* sampled_idx, const float* xyz, const int* feats,
                                     const float* centres, const float* values, float voxel_size, float* out) {
#pragma omp parallel for schedule(static)
  for (long long s = 0; s < M; ++s) {
    const int v = sampled_idx[s];
    float w[8], p[3];
    tri_weights(xyz + s * 3, centres + (long long)v * 3, voxel_size, w, p);
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
      for (int j = 0; j < 8; ++j) acc += w[j] * values[(long long)feats[(long long)v * 8 + j] * D + d];
      out[s * D + d] = acc;
    }
  }
}
/* backward of the above (what autograd derives): accumulated in double so that it can referee both
 * the reference's and the product's atomics-ordered float sums */
ORACLE_API void oracle_trilinear_bwd(long long M, int D, long long Kc, const int* sampled_idx, const float* xyz,
                                     const int* feats, const float* centres, const float* values, float voxel_size,
                                     const float* grad_out, float* grad_values, float* grad_xyz) {
  double* acc = (double*)calloc((size_t)Kc * D, sizeof(double));
  for (long long s = 0; s < M; ++s) {
    const int v = sampled_idx[s];
    float w[8], p[3];
    tri_weights(xyz + s * 3, centres + (long long)v * 3, voxel_size, w, p);
    double g3[3] = {0, 0, 0};
    for (int j = 0; j < 8; ++j) {
      const long long row = (long long)feats[(long long)v * 8 + j] * D;
      const int q[3] = {(j >> 2) & 1, (j >> 1) & 1, j & 1};
      double dot = 0;
      for (int d = 0; d < D; ++d) {
        acc[row + d] += (double)w[j] * grad_out[s * D + d];
        dot += (double)grad_out[s * D + d] * values[row + d];
      }
      double t[3];
      for (int a = 0; a < 3; ++a) t[a] = q[a] ? p[a] : 1.0 - p[a];
      g3[0] += (q[0] ? 1 : -1) * t[1] * t[2] * dot;
      g3[1] += (q[1] ? 1 : -1) * t[0] * t[2] * dot;
      g3[2] += (q[2] ? 1 : -1) * t[0] * t[1] * dot;
    }
    if (grad_xyz) for (int a = 0; a < 3; ++a) grad_xyz[s * 3 + a] = (float)(g3[a] / voxel_size);
  }
  for (long long i = 0; i < Kc * D; ++i) grad_values[i] = (float)acc[i];
  free(acc);
}

/* ---- compositing: VolumeRenderer.forward_chunk, fairnr/modules/renderer.py:193-218 ------------------ */
ORACLE_API void oracle_composite_fwd(long long B, int K, const float* fe, const float* tex, const float* depth,
                                     float* probs, float* out_depth, float* out_missed, float* out_colors) {
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < B; ++r) {
    float cum = 0.f; /* cumsum of the shifted free energy, float32 like torch.cumsum(x.float()) */
    float sp = 0.f, sd = 0.f, sc[3] = {0, 0, 0};
    for (int k = 0; k < K; ++k) {
      const float x = fe[r * K + k];
      const float a = 1.f - expf(-x);
      const float bb = expf(-cum);
      const float p = a * bb;
      probs[r * K + k] = p;
      sp += p;
      sd += depth[r * K + k] * p;
      if (tex) for (int c = 0; c < 3; ++c) sc[c] += tex[(r * K + k) * 3 + c] * p;
      cum += x;
    }
    out_depth[r] = sd;
    out_missed[r] = 1.f - sp;
    if (tex && out_colors) for (int c = 0; c < 3; ++c) out_colors[r * 3 + c] = sc[c];
  }
}
/* analytic backward in double (referee for autograd of the reference and for the fused kernel) */
ORACLE_API void oracle_composite_bwd(long long B, int K, const float* fe, const float* tex, const float* depth,
                                     const float* g_probs, const float* g_depth, const float* g_missed,
                                     const float* g_colors, float* g_fe, float* g_tex) {
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < B; ++r) {
    double* G = (double*)malloc(sizeof(double) * K * 3);
    double *P = G + K, *Bk = G + 2 * K;
    double cum = 0;
    for (int k = 0; k < K; ++k) {
      const double x = fe[r * K + k];
      Bk[k] = exp(-cum);
      P[k] = (1 - exp(-x)) * Bk[k];
      cum += x;
      double g = (g_probs ? g_probs[r * K + k] : 0) + (g_depth ? g_depth[r] * depth[r * K + k] : 0) -
                 (g_missed ? g_missed[r] : 0);
      if (tex && g_colors) for (int c = 0; c < 3; ++c) g += (double)g_colors[r * 3 + c] * tex[(r * K + k) * 3 + c];
      G[k] = g;
      if (g_tex) for (int c = 0; c < 3; ++c) g_tex[(r * K + k) * 3 + c] = (float)((g_colors ? g_colors[r * 3 + c] : 0) * P[k]);
    }
    double suffix = 0;
    for (int k = K - 1; k >= 0; --k) {
      const double x = fe[r * K + k];
      g_fe[r * K + k] = (float)(G[k] * exp(-x) * Bk[k] - suffix);
      suffix += G[k] * P[k];
    }
    free(G);
  }
}
