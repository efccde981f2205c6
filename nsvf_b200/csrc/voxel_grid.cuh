// Dense lattice of voxel indices + the ray walk over it (voxel_grid.cu); shared with aabb_intersect.cu, whose
// hierarchy kernels step aside when the lattice is usable.
#pragma once
#include "common.cuh"

namespace nsvf {

// 128 bytes of device memory in front of every voxel set's cell array.  Everything in it is produced on the device
// (no host round trip between "is this voxel set a lattice?" and the traversal that depends on the answer): the
// walk and the hierarchy kernels are both launched and exactly one of them returns at its first instruction.
struct VoxelGridHeader {
  int min_key[3];   // order-preserving integer image of the smallest voxel centre per axis
  int dims[3];      // lattice extent in cells
  int bad;          // != 0: centres off a common lattice, two voxels in one cell, non-finite centres
  int ok;           // extent fits the cell capacity (written once the extent is known)
  long long cap;    // cells available behind the header
  int pad[22];
};
static_assert(sizeof(VoxelGridHeader) == 128, "VoxelGridHeader is 128 bytes");

#ifdef __CUDACC__
__device__ __forceinline__ bool voxel_grid_usable(const unsigned char* ws, size_t per_set_bytes, int set) {
  if (ws == nullptr) return false;
  const VoxelGridHeader* h = reinterpret_cast<const VoxelGridHeader*>(ws + (size_t)set * per_set_bytes);
  return h->ok != 0 && h->bad == 0;
}
#endif

// Octree queries (svo_intersect.cu) walk the lattice of the LEAVES: `rank` (per node, DFS emission rank of a reachable
// leaf, -1 for everything else) selects the members of the voxel set and replaces the voxel index as tie-break /
// truncation key; veto[set] != 0 (the tree failed its checks) hands the set back to the traversal kernel;
// active[set] receives whether the walk took the set; defer[ray] = 1 marks rays left to the traversal kernel.
struct WalkOctree {
  const int* rank;
  long long rank_stride;
  const int* veto;
  int* active;
  unsigned char* defer;
};

// bytes of one voxel set's header + cells (multiple of 128); 0 when the lattice path is switched off
size_t voxel_grid_bytes(int n);
// filter (optional, per set `filter_stride` ints apart): points with filter[i] < 0 are not members of the set
int voxel_grid_build(cudaStream_t stream, int n_sets, int n, const float* points, long long points_stride,
                     float voxelsize, unsigned char* ws, size_t per_set_bytes, const int* filter = nullptr,
                     long long filter_stride = 0);
// mode as in aabb_intersect.cu: 1 = sorted by entry depth, 2 = any-hit mask
int voxel_grid_walk(cudaStream_t stream, int mode, const unsigned char* ws, size_t per_set_bytes, int n_sets, int n,
                    const float* points, long long points_stride, float voxelsize, long long rays_per_set, int n_max,
                    float empty_depth, const float* ray_start, const float* ray_dir, int* idx, float* min_depth,
                    float* max_depth, unsigned char* hit, const WalkOctree* octree = nullptr);

// sort_hits_by_depth (aabb_intersect.cu) restricted to the rays the walk deferred (walk_active / defer may be NULL: all rays)
int sort_hits_run(cudaStream_t stream, long long rays, int n_max, float empty_depth, int* idx, float* min_depth,
                  float* max_depth, unsigned char* hits, const int* walk_active, long long rays_per_tree,
                  const unsigned char* defer);

}  // namespace nsvf
