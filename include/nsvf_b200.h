/* nsvf_b200 — C ABI of the B200-native (sm_100a) sparse-voxel ray-marching hot path.
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (paths relative to the NSVF reference checkout).  Conventions:
 *   - all pointers are DEVICE pointers unless the parameter is documented as host;
 *   - tensors are dense, row-major, float32 / int32 exactly as the reference's _ext requires
 *     (fairnr/clib/include/utils.h:10-30 CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT);
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); every call is asynchronous on it,
 *     like the reference's at::cuda::getCurrentCUDAStream() launches (intersect_gpu.cu:367,
 *     sample_gpu.cu:209);
 *   - return value 0 = success; anything else is an error whose text nsvf_last_error() returns
 *     (thread-local).  The reference prints and calls exit(-1) (include/cuda_utils.h:35-44).
 *   - the library never allocates device memory: scratch is passed in as `workspace`
 *     (128-byte aligned, size from the matching *_workspace_bytes query).
 */
#ifndef NSVF_B200_H_
#define NSVF_B200_H_

#include <stddef.h>

#define NSVF_B200_VERSION 100

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NSVF_API __attribute__((visibility("default")))
#else
#define NSVF_API
#endif

typedef void* nsvf_stream_t;

NSVF_API int nsvf_version(void);
NSVF_API const char* nsvf_last_error(void);

/* y[i] = __fdividef(1.0f, x[i]) — the reciprocal the reference slab test uses
 * (fairnr/clib/src/intersect_gpu.cu:86-90). Test helper: lets a CPU oracle consume the exact values. */
NSVF_API int nsvf_ref_rcp(nsvf_stream_t stream, long long n, const float* x, float* y);

/* ---- ray / voxel intersection ------------------------------------------------------------------
 * Replaces aabb_intersect, fairnr/clib/src/intersect.cpp:49-75 + intersect_gpu.cu:125-167.
 *   ray_start, ray_dir : f32 [b, m, 3]
 *   points             : f32 [n, 3] shared by all batches when points_batch_stride == 0, otherwise
 *                        batch i reads points + i * points_batch_stride (floats) — the reference's
 *                        [b, n, 3] layout is points_batch_stride = 3 * n
 *   idx                : i32 [b, m, n_max]  first n_max hit voxels in ascending voxel index, then -1
 *   min_depth/max_depth: f32 [b, m, n_max]  entry / exit depth of each hit, 0 in unused slots
 * The outputs are fully written by the call (no pre-fill needed). */
NSVF_API size_t nsvf_aabb_workspace_bytes(int n, int n_trees /* 1 if points_batch_stride == 0 else b */);
NSVF_API int nsvf_aabb_intersect(nsvf_stream_t stream, int b, int n, int m, float voxelsize, int n_max,
                        const float* ray_start, const float* ray_dir, const float* points,
                        long long points_batch_stride, int* idx, float* min_depth, float* max_depth,
                        void* workspace, size_t workspace_bytes);

/* Replaces svo_intersect, fairnr/clib/src/intersect.cpp:84-112 + intersect_gpu.cu:170-237.
 *   points   : f32 [T, 3] node centres, children : i32 [T, 9] (slot 8 = node size in voxels, 1 = leaf),
 *              root = node T-1; shared by all batches when tree_batch_stride_nodes == 0, otherwise
 *              batch i reads node arrays offset by i * tree_batch_stride_nodes nodes
 *   idx      : leaf NODE indices in the reference's DFS emission order (children pushed 0..7,
 *              popped last-in-first-out), truncated at n_max; -1 / 0 fill as above.
 * The reference's per-ray stack bound (256, device assert) is kept: a ray that would exceed it stops
 * early and sets an internal flag instead of trapping. */
NSVF_API size_t nsvf_svo_workspace_bytes(int T, int n_trees);
NSVF_API int nsvf_svo_intersect(nsvf_stream_t stream, int b, int T, int m, float voxelsize, int n_max,
                       const float* ray_start, const float* ray_dir, const float* points, const int* children,
                       long long tree_batch_stride_nodes, int* idx, float* min_depth, float* max_depth,
                       void* workspace, size_t workspace_bytes);

/* ---- ray sampling ---------------------------------------------------------------------------------
 * Replaces inverse_cdf_sampling, fairnr/clib/src/sample.cpp:58-95 + sample_gpu.cu:108-202, and the
 * tiling / padding / noise / trimming glue of InverseCDFRaySampling.forward, fairnr/clib/__init__.py:231-300.
 *   pts_idx i32 / min_depth, max_depth, probs f32 : [b, num_rays, max_hits];  steps f32 [b, num_rays]
 *   uniform_noise f32 [b, num_rays, max_steps], or NULL to use `noise_const` everywhere (the wrapper's
 *   deterministic 0.5);  fixed_step_size <= 0 means 1 / steps
 *   sampled_idx i32 / sampled_depth, sampled_dists f32 : [valid_rays, max_steps], fully written
 *   (-1 / 0 / 0 beyond each ray's samples).
 *   valid_rays : only the first valid_rays of the b*num_rays rays exist in memory; the rest are the
 *                wrapper's padding copies of ray 0 (clib/__init__.py:237-242) and are neither stored nor
 *                computed (pass b*num_rays, or a negative value, for "all").
 *   ray_chunk  : the wrapper calls the reference kernel on column slices [:, i:i+ray_chunk] (:259-270);
 *                two reference quirks read a NEIGHBOURING ray (ray 0 of the block row, and slot 0 of the
 *                next ray in memory), so the slice width is part of the result. 0 = no slicing.
 *   max_count  : i32 [1] device, optional; atomicMax'ed with the largest number of samples (idx != -1)
 *                of any ray — the wrapper's max_len (:286). Zero it before the call. */
NSVF_API int nsvf_inverse_cdf_sampling(nsvf_stream_t stream, int b, int num_rays, long long valid_rays,
                                       int ray_chunk, int max_hits, int max_steps, float fixed_step_size,
                                       const int* pts_idx, const float* min_depth, const float* max_depth,
                                       const float* uniform_noise, float noise_const, const float* probs,
                                       const float* steps, int* sampled_idx, float* sampled_depth,
                                       float* sampled_dists, int* max_count);

/* Replaces uniform_ray_sampling, fairnr/clib/src/sample.cpp:23-55 + sample_gpu.cu:15-106. */
NSVF_API int nsvf_uniform_ray_sampling(nsvf_stream_t stream, int b, int num_rays, int max_hits, int max_steps,
                              float step_size, const int* pts_idx, const float* min_depth, const float* max_depth,
                              const float* uniform_noise, int* sampled_idx, float* sampled_depth,
                              float* sampled_dists);

/* ---- octree construction (HOST pointers, CPU) ---------------------------------------------------------
 * Replaces build_octree, fairnr/clib/src/octree.cpp:125-135.  Two calls on the same thread:
 *   nsvf_octree_build   : center f64[3], points i64 [n,3] (integer voxel coordinates), depth
 *                         -> total node count T and leaf count
 *   nsvf_octree_flatten : centers i32 [T,3], children i32 [T,9] */
NSVF_API int nsvf_octree_build(const double* center, const long long* points, long long n, int depth,
                      long long* out_total, long long* out_terminal);
NSVF_API int nsvf_octree_flatten(int* centers, int* children, long long capacity_nodes);

/* ---- trilinear voxel-corner embedding interpolation ------------------------------------------------
 * Replaces SparseVoxelEncoder.forward, fairnr/modules/encoder.py:582-590 with
 * trilinear_interp / offset_points, fairnr/data/geometry.py:195-200, 229-238, and its autograd backward.
 *   sampled_idx i32 [M] voxel of each sample (>= 0), sampled_xyz f32 [M,3]
 *   feats i32 [n,8] corner keys per voxel (corner order x slowest, z fastest), centres f32 [n,3]
 *   values f32 [Kc, D];  out f32 [M, D]
 * bwd: grad_values f32 [Kc, D] is ACCUMULATED into (zero it first); grad_xyz f32 [M,3] or NULL. */
NSVF_API int nsvf_trilinear_embed_fwd(nsvf_stream_t stream, long long M, int D, const int* sampled_idx,
                             const float* sampled_xyz, const int* feats, const float* centres,
                             const float* values, float voxel_size, float* out);
NSVF_API int nsvf_trilinear_embed_bwd(nsvf_stream_t stream, long long M, int D, const int* sampled_idx,
                             const float* sampled_xyz, const int* feats, const float* centres,
                             const float* values, float voxel_size, const float* grad_out, float* grad_values,
                             float* grad_xyz);

/* ---- alpha compositing -------------------------------------------------------------------------------
 * Replaces the compositing block of VolumeRenderer.forward_chunk, fairnr/modules/renderer.py:193-218.
 *   free_energy f32 [B,K] (0 at invalid samples), texture f32 [B,K,3] or NULL, sampled_depth f32 [B,K]
 *   -> probs f32 [B,K] (or NULL), depth f32 [B], missed f32 [B], colors f32 [B,3] (or NULL)
 * bwd: any grad_* input may be NULL (treated as zero); grad_texture may be NULL. */
NSVF_API int nsvf_composite_fwd(nsvf_stream_t stream, long long B, int K, const float* free_energy, const float* texture,
                       const float* sampled_depth, float* probs, float* depth, float* missed, float* colors);
NSVF_API int nsvf_composite_bwd(nsvf_stream_t stream, long long B, int K, const float* free_energy, const float* texture,
                       const float* sampled_depth, const float* grad_probs, const float* grad_depth,
                       const float* grad_missed, const float* grad_colors, float* grad_free_energy,
                       float* grad_texture);

#ifdef __cplusplus
}
#endif
#endif /* NSVF_B200_H_ */
