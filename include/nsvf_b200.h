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

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
NSVF_API unsigned long long nsvf_kernel_launches(void);

/* Measurement hook: cudaEvent_t `ev_start` / `ev_stop` are recorded on the launching stream immediately
 * before / after every launch of the kernel called `name` (e.g. "aabb_intersect_kernel"), so a caller can
 * time one kernel inside a longer step without a profiler.  name = NULL or "" clears the hook. */
NSVF_API int nsvf_profile_kernel(const char* name, void* ev_start, void* ev_stop);
/* Accumulating variant: between nsvf_profile_begin(name) and nsvf_profile_end every launch of kernel `name` is bracketed
 * by its own event pair (pool of 8192 pairs owned by the library, recorded on the launching stream);
 * nsvf_profile_end waits for the recorded events and returns the number of launches, their summed device time and the
 * shortest / longest launch (ms).  This is how bench.py times a kernel "live", inside the step it belongs to. */
NSVF_API int nsvf_profile_begin(const char* name);
NSVF_API int nsvf_profile_end(int* n_launches, float* total_ms, float* min_ms, float* max_ms);

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
 * The outputs are fully written by the call (no pre-fill needed).
 * How: voxel centres that lie on a regular lattice (every NSVF voxel set does) are scattered into a dense cell array
 * and each ray walks the cells it passes through, running the reference's slab test on the occupied ones
 * (csrc/voxel_grid.cu); any other point set goes through an 8-ary hierarchy of enclosing boxes
 * (csrc/aabb_intersect.cu).  The choice is made on the device, per voxel set, and does not change results.
 * Workspace: hierarchy + 16 cells (4 B each) per voxel. */
NSVF_API size_t nsvf_aabb_workspace_bytes(int n, int n_trees /* 1 if points_batch_stride == 0 else b */);
NSVF_API int nsvf_aabb_intersect(nsvf_stream_t stream, int b, int n, int m, float voxelsize, int n_max,
                        const float* ray_start, const float* ray_dir, const float* points,
                        long long points_batch_stride, int* idx, float* min_depth, float* max_depth,
                        void* workspace, size_t workspace_bytes);

/* Same intersection with the post-processing of SparseVoxelEncoder.ray_intersect fused in
 * (fairnr/modules/encoder.py:519-524: masked_fill(idx == -1, MAX_DEPTH), sort by min_depth, gather, any):
 * each ray's hits are sorted by entry depth (ties: ascending voxel index), unused slots hold idx -1 and
 * `empty_depth` (the encoder uses 10000.0), and hits u8 [b, m] (optional) = "the ray hit something". */
NSVF_API int nsvf_aabb_intersect_sorted(nsvf_stream_t stream, int b, int n, int m, float voxelsize, int n_max,
                                        float empty_depth, const float* ray_start, const float* ray_dir,
                                        const float* points, long long points_batch_stride, int* idx,
                                        float* min_depth, float* max_depth, unsigned char* hits, void* workspace,
                                        size_t workspace_bytes);

/* The same post-processing as a stand-alone, in-place pass over hit lists produced by any intersection routine
 * (used after nsvf_svo_intersect): idx i32 / min_depth, max_depth f32 [rays, n_max] are rewritten sorted by entry
 * depth with -1 / empty_depth in the unused slots; hits u8 [rays] optional. */
NSVF_API int nsvf_sort_hits_by_depth(nsvf_stream_t stream, long long rays, int n_max, float empty_depth, int* idx,
                                     float* min_depth, float* max_depth, unsigned char* hits);

/* Any-hit query: hits u8 [b, m] = 1 iff nsvf_aabb_intersect would report at least one hit for the ray.  Lets
 * the training path (`--no-sampling-at-reader`, fairnr/models/nsvf.py:48-60) draw pixels from the hit mask of
 * all V x H x W rays without materialising [rays, n_max] x 3 outputs for rays it will not march. */
NSVF_API int nsvf_aabb_hit_mask(nsvf_stream_t stream, int b, int n, int m, float voxelsize, const float* ray_start,
                                const float* ray_dir, const float* points, long long points_batch_stride,
                                unsigned char* hits, void* workspace, size_t workspace_bytes);

/* The voxel set only changes when voxels are pruned or split (fairnr/modules/encoder.py:440-505), the rays change every
 * call: nsvf_aabb_prepare fills a workspace of nsvf_aabb_workspace_bytes(n, n_sets) once (n_sets = 1 with
 * points_batch_stride 0, else one set per batch row) and nsvf_aabb_intersect_prepared runs any of the three queries
 * above on it without rebuilding — mode 0: nsvf_aabb_intersect (empty_depth ignored, hits may be NULL),
 * 1: nsvf_aabb_intersect_sorted, 2: nsvf_aabb_hit_mask (idx / depths may be NULL).  `points` must be the centres the
 * workspace was prepared from (the exact test reads them).  The one-shot entry points are prepare + this. */
NSVF_API int nsvf_aabb_prepare(nsvf_stream_t stream, int n_sets, int n, float voxelsize, const float* points,
                               long long points_batch_stride, void* workspace, size_t workspace_bytes);
NSVF_API int nsvf_aabb_intersect_prepared(nsvf_stream_t stream, int mode, int b, int n, int m, float voxelsize,
                                          int n_max, float empty_depth, const float* ray_start, const float* ray_dir,
                                          const float* points, long long points_batch_stride, int* idx,
                                          float* min_depth, float* max_depth, unsigned char* hits,
                                          const void* workspace, size_t workspace_bytes);

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

/* nsvf_svo_intersect followed by nsvf_sort_hits_by_depth (what SparseVoxelEncoder.ray_intersect does on the octree path,
 * fairnr/modules/encoder.py:495-524) as ONE call with the same outputs: idx / depths sorted by entry depth (ties: DFS
 * order), -1 / empty_depth fill, hits u8 [b, m] optional.  When the tree passes its consistency checks and the leaves
 * lie on a lattice the rays walk that lattice instead of descending the tree (csrc/svo_intersect.cu, bottom); the
 * results are the same either way.  Workspace: nsvf_svo_sorted_workspace_bytes(T, n_trees, b * m). */
NSVF_API size_t nsvf_svo_sorted_workspace_bytes(int T, int n_trees, long long rays);
NSVF_API int nsvf_svo_intersect_sorted(nsvf_stream_t stream, int b, int T, int m, float voxelsize, int n_max,
                                       float empty_depth, const float* ray_start, const float* ray_dir,
                                       const float* points, const int* children, long long tree_batch_stride_nodes,
                                       int* idx, float* min_depth, float* max_depth, unsigned char* hits,
                                       void* workspace, size_t workspace_bytes);
/* The octree only changes when voxels are pruned or split: nsvf_svo_prepare fills a workspace of
 * nsvf_svo_sorted_workspace_bytes(T, n_trees, 0) once (packed nodes, consistency checks, DFS ranks, leaf lattice) and
 * nsvf_svo_intersect_sorted_prepared answers rays on it; ray_scratch: nsvf_svo_ray_scratch_bytes(n_trees, b * m) bytes,
 * 128-byte aligned, any contents.  points / children must be the arrays the workspace was prepared from. */
NSVF_API int nsvf_svo_prepare(nsvf_stream_t stream, int n_trees, int T, float voxelsize, const float* points,
                              const int* children, long long tree_batch_stride_nodes, void* workspace,
                              size_t workspace_bytes);
NSVF_API size_t nsvf_svo_ray_scratch_bytes(int n_trees, long long rays);
NSVF_API int nsvf_svo_intersect_sorted_prepared(nsvf_stream_t stream, int b, int T, int m, float voxelsize, int n_max,
                                                float empty_depth, const float* ray_start, const float* ray_dir,
                                                const float* points, const int* children,
                                                long long tree_batch_stride_nodes, int* idx, float* min_depth,
                                                float* max_depth, unsigned char* hits, const void* workspace,
                                                size_t workspace_bytes, void* ray_scratch, size_t ray_scratch_bytes);

/* The two clib entry points outside the NSVF path (kept for API completeness of the 7-function module):
 * ball_intersect, fairnr/clib/src/intersect.cpp:15-44 + intersect_gpu.cu:15-70 (no caller in the reference), and
 * triangle_intersect, intersect.cpp:120-146 + intersect_gpu.cu:240-347 (mesh encoder):
 *   face_points f32 [b, n, 9]; depth f32 [b, m, n_max*3] = (t, -cage_near, cage_far) per hit sorted by t;
 *   uv f32 [b, m, n_max*2].  Batch strides as for nsvf_aabb_intersect (0 = shared). Outputs fully written. */
NSVF_API int nsvf_ball_intersect(nsvf_stream_t stream, int b, int n, int m, float radius, int n_max,
                                 const float* ray_start, const float* ray_dir, const float* points,
                                 long long points_batch_stride, int* idx, float* min_depth, float* max_depth);
NSVF_API int nsvf_triangle_intersect(nsvf_stream_t stream, int b, int n, int m, float cagesize, float blur, int n_max,
                                     const float* ray_start, const float* ray_dir, const float* face_points,
                                     long long faces_batch_stride, int* idx, float* depth, float* uv);

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

/* The same sampler with the post-processing of SparseVoxelEncoder.ray_sample (fairnr/modules/encoder.py:547-549:
 * dists.clamp(min=0), depth[idx == -1] = MAX_DEPTH, dists[idx == -1] = 0) fused in and TRIMMED rows:
 *   ray_len    : i32 [valid_rays], optional: 1 + position of the last sample with idx != -1 of each ray
 *   holes_flag : i32 [1], optional, OR'ed with 1 when some ray's valid samples are not a prefix of its row
 *   pad_depth  : depth written to padding slots (0 reproduces the reference kernel, 10000 = MAX_DEPTH for ray_sample)
 *   flags      : bit 0 = write the padding beyond each ray's samples (idx -1, depth pad_depth, dists 0); without it
 *                the rows hold exactly the produced samples and nothing else is written (consumers use ray_len);
 *                bit 1 = apply ray_sample's clamp / masking to the produced samples. */
NSVF_API int nsvf_inverse_cdf_sampling_ex(nsvf_stream_t stream, int b, int num_rays, long long valid_rays,
                                          int ray_chunk, int max_hits, int max_steps, float fixed_step_size,
                                          const int* pts_idx, const float* min_depth, const float* max_depth,
                                          const float* uniform_noise, float noise_const, const float* probs,
                                          const float* steps, int* sampled_idx, float* sampled_depth,
                                          float* sampled_dists, int* max_count, int* ray_len, int* holes_flag,
                                          float pad_depth, int flags);

/* On-demand variant for the ray-marching plan (with early termination most emitted samples are never evaluated):
 *   nsvf_inverse_cdf_plan  : per ray, ONLY the number of samples nsvf_inverse_cdf_sampling_ex would emit
 *                            (ray_len i32 [valid_rays]) in O(bins + log steps), plus the two neighbour-ray values the
 *                            reference's trailing loop reads (quirk i32 [valid_rays, 2] = {slot 0 of the next ray,
 *                            valid bins nb0 of ray 0 of the block row, stored as -1 - nb0 for a ray with irregular
 *                            cumulative sums});
 *                            meta i32 [3] (zeroed) = {max ray_len, holes flag, fallback rays}.  Hit lists must be
 *                            -1-terminated (what the sorted intersection produces).
 *   nsvf_inverse_cdf_block : the samples at positions [k_begin, k_end) of the rays not flagged in early_stop, with
 *                            ray_sample's post-processing applied, written into the slot-major planes idxT / depthT /
 *                            distsT [K][nsvf_march_plane_stride(B)] of the plan.  The per-ray inputs may be row slices of
 *                            the tensors the plan kernel saw (uniform_noise rows noise_row_stride floats apart, or NULL).
 *                            Any block, in any order (each block rebuilds the ray's tables).
 *   nsvf_inverse_cdf_stream: the same output for blocks requested IN INCREASING ORDER — k_begin = 0 first, then every
 *                            call's k_begin equal to the previous call's k_end — which is how the ray-marching loop
 *                            asks: one thread per ray resumes the reference's serial state machine from `state`
 *                            (nsvf_inverse_cdf_stream_state_bytes(B) bytes, 16-byte aligned, owned by the caller between
 *                            calls; rays flagged in early_stop are skipped and must stay flagged).
 * Same samples, bit for bit, as the eager kernel + nsvf_march_transpose. */
NSVF_API int nsvf_inverse_cdf_plan(nsvf_stream_t stream, int b, int num_rays, long long valid_rays, int ray_chunk,
                                   int max_hits, int max_steps, float fixed_step_size, const int* pts_idx,
                                   const float* min_depth, const float* max_depth, const float* uniform_noise,
                                   float noise_const, const float* probs, const float* steps, int* ray_len, int* quirk,
                                   int* meta);
NSVF_API int nsvf_inverse_cdf_block(nsvf_stream_t stream, long long B, int max_hits, int max_steps,
                                    float fixed_step_size, int k_begin, int k_end, const unsigned char* early_stop,
                                    const int* ray_len, const int* quirk, const int* pts_idx, const float* min_depth,
                                    const float* max_depth, const float* uniform_noise, long long noise_row_stride,
                                    float noise_const, const float* probs, const float* steps, float pad_depth,
                                    int* idxT, float* depthT, float* distsT);

NSVF_API size_t nsvf_inverse_cdf_stream_state_bytes(long long B);
NSVF_API int nsvf_inverse_cdf_stream(nsvf_stream_t stream, long long B, int max_hits, int max_steps,
                                     float fixed_step_size, int k_begin, int k_end, const unsigned char* early_stop,
                                     const int* ray_len, const int* quirk, const int* pts_idx, const float* min_depth,
                                     const float* max_depth, const float* uniform_noise, long long noise_row_stride,
                                     float noise_const, const float* probs, const float* steps, float pad_depth,
                                     void* state, int* idxT, float* depthT, float* distsT);

/* Replaces uniform_ray_sampling, fairnr/clib/src/sample.cpp:23-55 + sample_gpu.cu:15-106, and the trimming glue of
 * UniformRaySampling.forward, fairnr/clib/__init__.py:178-228.  Same layouts as above ([b, num_rays, max_hits] bins,
 * noise / outputs [b, num_rays, max_steps]); outputs fully written: beyond a ray's samples idx -1, depth 0, dists 0
 * (the reference leaves stale values of its in-place merge there).  max_count i32 [1] (optional, zero it first) is
 * atomicMax'ed with the largest number of samples of any ray — the wrapper's max_len (:214). */
NSVF_API int nsvf_uniform_ray_sampling(nsvf_stream_t stream, int b, int num_rays, int max_hits, int max_steps,
                              float step_size, const int* pts_idx, const float* min_depth, const float* max_depth,
                              const float* uniform_noise, int* sampled_idx, float* sampled_depth,
                              float* sampled_dists, int* max_count);

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

/* ---- ray-marching plan: the chunk loop of the renderer on the device ---------------------------------------
 * Replaces the loop of VolumeRenderer.forward_chunk / forward_once, fairnr/modules/renderer.py:77-191: per-column
 * `hits[:, i].sum()` host syncs (:157-158,187), boolean-mask compaction (:88-100), masked_scatter into zero-filled
 * [B,K] tensors (:109-131), free_energy = relu(noise + sigma) * dists * 7 (:117-121), the early-termination
 * update after every field evaluation (:170-174), and the compositing block (:193-218) with its two per-ray depth
 * extrema (:210-211).  Same schedule, same samples reach the field.
 * Input rows are TRIMMED: ray r owns slots [0, lens[r]) of row r of sampled_idx / sampled_depth / sampled_dists (row
 * stride ldk >= K); valid samples must be a prefix of the row (nsvf_march_ray_lengths reports rows where they are
 * not in plan word 4; callers then use nsvf_compact_* below).  Inside the plan everything is SLOT-MAJOR: planes
 * [K][ldb] with ldb = nsvf_march_plane_stride(B) >= B (entry (k, r) at k*ldb + r, defined for k < lens[r]; texture
 * planes [K][ldb][3]), so that one thread per ray walks a column window with coalesced accesses.
 *   plan        : device scratch of nsvf_march_plan_bytes(B, K) bytes, ZEROED once per forward_chunk call
 *   host_info   : i32 [>= 16 (+ 3 per window)] in PINNED host memory (written by the device through the unified
 *                 address space): [0] start, [1] end, [2] valid samples of the window, [3] done, [4] holes,
 *                 [5] number of windows, [8] total samples; with all_windows the triplets (start, end, count) of
 *                 every window follow from word 16.  Read it after synchronising the stream.
 *   nsvf_march_ray_lengths : lens i32 [B] from a padded sampled_idx (idx != -1)
 *   nsvf_march_transpose   : columns [k_begin, k_end) (k_begin a multiple of 32) of idxT i32 / depthT, distsT f32 from
 *                            the rows (tiled shared-memory transpose); rays flagged in early_stop (optional) are
 *                            skipped, so with early termination the planes are filled block by block as the window
 *                            loop advances and stopped rays cost nothing
 *   nsvf_march_begin       : column counts from lens; publishes the first window, or (all_windows = 1, valid when
 *                            there is no early termination) the complete window list
 *   nsvf_march_compact     : compacts the samples of window [start, end) of the live rays in row-major order (the
 *                            order boolean indexing produces): out_vox i32 [M], out_xyz f32 [M,3] = ray_start +
 *                            ray_dir * depth, out_dir f32 [M,3], out_dists f32 [M] (either optional), and
 *                            ray_off i32 [B+1] = exclusive offsets of the rays in that order.  launch_no: ignored
 *                            (kept for callers of the first-generation scan; pass 0).  start = -1: the window
 *                            is the one the preceding nsvf_march_epilogue (schedule_next) left in the plan, so the
 *                            launch can be queued BEFORE the host has read that window back; the outputs must then
 *                            hold max(chunk_size, B) rows (nothing is written if the schedule is finished).
 *   nsvf_march_epilogue    : sigma f32 [M] (+ noise f32 [M] or NULL, dists f32 [M]) -> feT f32 [K][B],
 *                            texture f32 [M,3] -> texT f32 [K][B][3] at the window's slots; eval_len i32 [B] =
 *                            evaluated prefix; with tolerance > 0: acc_free_energy f32 [B] += row sums, early_stop
 *                            u8 [B] = acc > tolerance, column counts updated; with schedule_next the next window is
 *                            published to plan / host_info by the last CTA.
 *   nsvf_march_epilogue_bwd: gradients of the planes back to the compacted order of one window:
 *                            grad_sigma = (g_fe * 7) * dists * [noise + sigma > 0], grad_texture = g_tex.
 *   nsvf_march_composite_fwd: probsT f32 [K][B] (optional, zeros beyond the evaluated prefix), depth / missed f32 [B],
 *                            colors f32 [B,3]; max_depths f32 [B] = max sample depth, -1 for rays flagged in
 *                            early_stop; min_depths f32 [B] = min over the WHOLE padded row: the samples and, when
 *                            lens[r] < K, the padding — the first padding slot of padded_depth_rows (row stride ldk)
 *                            when given, else pad_depth.  lazy_planes != 0: the planes of a ray flagged in early_stop
 *                            hold its evaluated prefix only (block-wise transpose); its minimum is then taken over
 *                            that prefix, which is exact for depth-ordered (ray-marched) samples.
 *   nsvf_march_composite_bwd: g_feT [K][B], g_texT [K][B][3] for the evaluated prefix; scratchT f32 [K][B]. */
NSVF_API long long nsvf_march_plane_stride(long long B);
NSVF_API size_t nsvf_march_plan_bytes(long long B, int K);
NSVF_API int nsvf_march_ray_lengths(nsvf_stream_t stream, long long B, int K, long long ldk, const int* sampled_idx,
                                    int* lens, void* plan);
NSVF_API int nsvf_march_transpose(nsvf_stream_t stream, long long B, int K, long long ldk, int k_begin, int k_end,
                                  const unsigned char* early_stop, const int* lens, const int* sampled_idx, const float* sampled_depth, const float* sampled_dists,
                                  int* idxT, float* depthT, float* distsT);
NSVF_API int nsvf_march_begin(nsvf_stream_t stream, long long B, int K, int chunk_size, const int* lens,
                              const unsigned char* early_stop, int all_windows, void* plan, int* host_info,
                              int host_capacity_ints);
NSVF_API int nsvf_march_compact(nsvf_stream_t stream, long long B, int K, int start, int end, const int* lens,
                                const unsigned char* early_stop, const int* idxT, const float* depthT,
                                const float* distsT, const float* ray_start, const float* ray_dir, int* out_vox,
                                float* out_xyz, float* out_dir, float* out_dists, int* ray_off, void* plan,
                                int launch_no);
NSVF_API int nsvf_march_epilogue(nsvf_stream_t stream, long long B, int K, int start, int end, const int* ray_off,
                                 const int* lens, unsigned char* early_stop, float* acc_free_energy, int* eval_len,
                                 const float* sigma, const float* noise, const float* dists, const float* texture,
                                 float tolerance, float* feT, float* texT, int chunk_size, int schedule_next,
                                 void* plan, int* host_info);
NSVF_API int nsvf_march_epilogue_bwd(nsvf_stream_t stream, long long B, int K, int start, int end, const int* ray_off,
                                     const float* g_feT, const float* g_texT, const float* sigma, const float* noise,
                                     const float* dists, float* grad_sigma, float* grad_texture);
NSVF_API int nsvf_march_composite_fwd(nsvf_stream_t stream, long long B, int K, const int* eval_len, const int* lens,
                                      const unsigned char* early_stop, const float* feT, const float* texT,
                                      const float* depthT, float* probsT, float* depth, float* missed, float* colors,
                                      float* max_depths, float* min_depths, const float* padded_depth_rows,
                                      long long ldk, float pad_depth, int lazy_planes);
NSVF_API int nsvf_march_composite_bwd(nsvf_stream_t stream, long long B, int K, const int* eval_len, const float* feT,
                                      const float* texT, const float* depthT, const float* grad_probsT,
                                      const float* grad_depth, const float* grad_missed, const float* grad_colors,
                                      float* g_feT, float* g_texT, float* scratchT);

/* ---- sample compaction ---------------------------------------------------------------------------------
 * Replaces the boolean-mask compaction of VolumeRenderer.forward_once, fairnr/modules/renderer.py:88-100
 * (sample_mask = idx != -1 & ~early_stop; xyz = ray_start + ray_dir * depth; s[sample_mask] for every tensor).
 *   sampled_idx i32 / sampled_depth, sampled_dists f32 : [B,K]; only columns [col0, col1) are considered
 *   early_stop u8 [B] or NULL (rays with a non-zero flag contribute no samples)
 *   nsvf_compact_count : counts i64 [B] = valid samples per ray
 *   nsvf_compact_fill  : offsets_incl i64 [B] = inclusive prefix sum of counts (any device scan);
 *                        writes, in row-major order of the valid samples, out_vox i32 [M], out_xyz f32 [M,3],
 *                        out_dir f32 [M,3] (or NULL), out_dists f32 [M] (or NULL), out_flat i64 [M] = ray*K + k */
NSVF_API int nsvf_compact_count(nsvf_stream_t stream, long long B, int K, int col0, int col1, const int* sampled_idx,
                                const unsigned char* early_stop, long long* counts);
NSVF_API int nsvf_compact_fill(nsvf_stream_t stream, long long B, int K, int col0, int col1, const int* sampled_idx,
                               const float* sampled_depth, const float* sampled_dists,
                               const unsigned char* early_stop, const float* ray_start, const float* ray_dir,
                               const long long* offsets_incl, int* out_vox, float* out_xyz, float* out_dir,
                               float* out_dists, long long* out_flat);

/* counts i32 [col1 - col0] (fully written): counts[k - col0] = number of rays with sampled_idx[ray,k] != -1 and
 * early_stop[ray] == 0, k in [col0, col1) — the per-column totals VolumeRenderer.forward_chunk's chunk scheduler
 * needs (fairnr/modules/renderer.py:158,187: one `hits[:, i].sum()` host sync per column in the reference). */
NSVF_API int nsvf_masked_col_counts(nsvf_stream_t stream, long long B, int K, int col0, int col1,
                                    const int* sampled_idx, const unsigned char* early_stop, int* counts);

/* ---- half-voxel splitting --------------------------------------------------------------------------------
 * Replaces splitting_points, fairnr/data/geometry.py:250-274 (+ discretize_points :241-247, offset_points
 * :229-238), called by SparseVoxelEncoder.splitting, fairnr/modules/encoder.py:656-676.
 * Two phases because the number of unique new corner keys Kc' sizes the outputs:
 *   pmin f32[3], max_coord i32[3] : HOST values — per-axis min of `points`, and per-axis max of
 *                                   round((points - pmin) / quarter_voxel), quarter_voxel = half_voxel / 2
 *   nsvf_split_mark : new_points f32 [8n,3] (children centres, child order x slowest - z fastest) and
 *                     *n_keys (device i32) = Kc'
 *   nsvf_split_emit : new_feats i32 [8n,8] = lexicographic rank of each child's corner keys (== torch.unique's
 *                     inverse), new_keys i32 [Kc',3] (optional), new_values f32 [Kc',D] (optional) = the parent's
 *                     trilinear interpolant at the key, parent = the smallest voxel index touching the key;
 *                     parent i32 [Kc'] is scratch.  `workspace` must be the one phase 1 filled. */
NSVF_API size_t nsvf_split_workspace_bytes(const int* max_coord);
NSVF_API int nsvf_split_mark(nsvf_stream_t stream, int n, const float* points, float half_voxel, const float* pmin,
                             const int* max_coord, float* new_points, int* n_keys, void* workspace,
                             size_t workspace_bytes);
NSVF_API int nsvf_split_emit(nsvf_stream_t stream, int n, int D, const float* points, const int* feats,
                             const float* values, float half_voxel, const float* pmin, const int* max_coord,
                             int n_keys, int* new_feats, int* parent, int* new_keys, float* new_values,
                             void* workspace, size_t workspace_bytes);

/* ---- pruning --------------------------------------------------------------------------------------------
 * Replaces SparseVoxelEncoder.get_scores / pruning, fairnr/modules/encoder.py:605-654.
 *   nsvf_prune_lattice_embed : out f32 [nv, bits^3, D] = interpolated embeddings at the bits^3 lattice points
 *                              (offset_points(points, voxel_size/2, bits), closed voxel) of voxels [v0, v0+nv);
 *                              feats i32 [n,8], centres f32 [n,3], values f32 [Kc,D]
 *   nsvf_prune_keep          : sigma f32 [nv, L] (the field's density at those points) ->
 *                              keep u8 [nv] = (1 - min_l exp(-relu(sigma))) > th; min_score f32 [nv] optional */
NSVF_API int nsvf_prune_lattice_embed(nsvf_stream_t stream, int nv, int v0, int bits, int D, const int* feats,
                                      const float* centres, const float* values, float voxel_size, float* out);
NSVF_API int nsvf_prune_keep(nsvf_stream_t stream, int nv, int L, const float* sigma, float th, unsigned char* keep,
                             float* min_score);

/* ---- per-frame post-processing ------------------------------------------------------------------------------
 * nsvf_fill_in_blend: replaces fill_in (fairnr/data/geometry.py:303-317) for colors / missed / depths and the
 * background blend of NSVFModel.postprocessing (fairnr/models/nsvf.py:89-104).
 *   hits u8 [N]; rank_incl i64 [N] = inclusive prefix sum of hits (row of the hit ray in the compacted results);
 *   colors f32 [M,3], missed f32 [M], depths f32 [M] (M = number of hit rays); bg_color f32 [3] (device)
 *   -> out_colors [N,3] = colors_full + missed_full * bg_color, out_missed [N] (1 where not hit),
 *      out_depths [N] = depths_full + missed_full * bg_depth
 * nsvf_track_voxel_probs: replaces SparseVoxelEncoder.track_voxel_probs (fairnr/modules/encoder.py:594-603):
 *   max_probs f32 [n_vox] (>= 0, updated in place) = max(max_probs, per-ray sums of probs per voxel).
 *   Precondition (true for ray-marched samples: they are depth-ordered and voxels are convex): the samples a ray
 *   has in one voxel are consecutive; the reference's scatter_add would also merge non-consecutive repeats. */
NSVF_API int nsvf_fill_in_blend(nsvf_stream_t stream, long long N, const unsigned char* hits,
                                const long long* rank_incl, const float* colors, const float* missed,
                                const float* depths, const float* bg_color, float bg_depth, float* out_colors,
                                float* out_missed, float* out_depths);
NSVF_API int nsvf_track_voxel_probs(nsvf_stream_t stream, long long B, int K, const int* sampled_idx,
                                    const float* probs, int n_vox, float* max_probs);

/* ---- field MLP: the passes around the Linear of an FCLayer ----------------------------------------------------
 * Replaces, for FCLayer (fairnr/modules/module_utils.py:97-111: nn.Linear -> nn.LayerNorm([N]) -> nn.ReLU), the
 * LayerNorm + ReLU forward and — in one pass — everything of its backward that is not a GEMM: the ReLU mask, d gamma,
 * d beta, the layer-norm input gradient dh and the Linear's bias gradient (column sums of dh).  The Linear's
 * contractions stay on cuBLAS.  N must be 128, 256 or 512; h, y, dy, dh f32 [M, N]; gamma, beta f32 [N];
 * mean, rstd f32 [M] (biased variance, rstd = rsqrt(var + eps), as at::native::layer_norm).
 *   fwd: y = max((h - mean) * rstd * gamma + beta, 0); mean / rstd (optional, needed by bwd)
 *   bwd: dh f32 [M, N]; dgamma, dbeta, dbias f32 [N] (each optional) — deterministic (fixed summation order). */
NSVF_API int nsvf_ln_relu_fwd(nsvf_stream_t stream, long long M, int N, const float* h, const float* gamma,
                              const float* beta, float eps, float* y, float* mean, float* rstd);
NSVF_API size_t nsvf_ln_relu_bwd_workspace_bytes(long long M, int N);
NSVF_API int nsvf_ln_relu_bwd(nsvf_stream_t stream, long long M, int N, const float* h, const float* dy,
                              const float* gamma, const float* beta, const float* mean, const float* rstd, float* dh,
                              float* dgamma, float* dbeta, float* dbias, void* workspace, size_t workspace_bytes);

/* NeRF positional encoding of the field inputs, NeRFPosEmbLinear(no_linear=True), fairnr/modules/module_utils.py:56-87:
 *   x f32 [M, C]; freq f32 [L] (device; L = 4, 6 or 10); angular: t = acos(clamp(x, -1+1e-6, 1-1e-6)) else t = x;
 *   out f32 [M, C*2L (+ C if cat_input)]: per channel c the 2L values sin(f_k t) (k < L) then cos(f_k t), then the C raw
 *   inputs (module_utils.py:80-86: outer product, cat([sin, cos], -1), view, cat([x, inputs], -1)).
 *   bwd (non-angular only: ray directions need no gradient): grad_x f32 [M, C]. */
NSVF_API int nsvf_posenc_fwd(nsvf_stream_t stream, long long M, int C, int L, const float* x, const float* freq,
                             int angular, int cat_input, float* out);
NSVF_API int nsvf_posenc_bwd(nsvf_stream_t stream, long long M, int C, int L, const float* x, const float* freq,
                             int cat_input, const float* grad_out, float* grad_x);

/* The output heads of the field (nn.Linear(128, 1) for sigma, nn.Linear(256, 3) for rgb; fairnr/modules/field.py via
 * FCBlock's outermost Linear, module_utils.py:114-150): y[M, O] = x[M, K] W[O, K]^T + b with O <= 4 output features.
 * Supported (K, O): see nsvf_narrow_linear_supported.  bwd: dx f32 [M, K] (optional), dW f32 [O, K], db f32 [O]
 * (each optional), deterministic. */
NSVF_API int nsvf_narrow_linear_supported(int K, int O);
NSVF_API int nsvf_narrow_linear_fwd(nsvf_stream_t stream, long long M, int K, int O, const float* x, const float* W,
                                    const float* b, float* y);
NSVF_API size_t nsvf_narrow_linear_bwd_workspace_bytes(long long M, int K, int O);
NSVF_API int nsvf_narrow_linear_bwd(nsvf_stream_t stream, long long M, int K, int O, const float* x, const float* W,
                                    const float* dy, float* dx, float* dW, float* db, void* workspace,
                                    size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* NSVF_B200_H_ */
