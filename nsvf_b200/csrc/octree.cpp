// build_octree: flattened sparse voxel octree (centers int32[T,3], children int32[T,9]).
//
// Replaces fairnr/clib/src/octree.cpp:13-135 (EasyOctree), which inserts points with one at::Tensor op
// and one .item<int>() per point per level (a device sync each when the tensors live on the GPU).
// Same numbering contract, plain arrays:
//   * child slot of a point below a node = (x > cx) + 2 (y > cy) + 4 (z > cz)           (:62-64)
//   * child centre = parent centre + (2 bit - 1) * 2^(depth-1) for depth >= 1           (:69-71)
//   * at depth 0 the child is the leaf itself (it REPLACES a previous leaf in that slot)   (:65-67)
//   * numbering: leaves keep their insertion index; internal nodes are numbered downwards from
//     T-1 (root) in BFS x slot order; children[k][8] = 1 << (depth+1) (leaf: depth -1 -> 1);
//     centers[k] = node centre truncated to int32                                       (:93-123)
// Centres are kept in double here; the reference keeps them in a float32/int64 tensor whose values are
// integers (or, for an odd-sum root, x.5 values that are exact in both) well below 2^24.
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "common.cuh"
#include "nsvf_b200.h"

namespace {

struct Node {
  double c[3];
  int depth;     // -1 = leaf
  int index;
  int child[8];  // index into the pool, -1 = none
};

struct Tree {
  std::vector<Node> pool;
  int new_node(const double* c, int depth, int index) {
    Node nd;
    nd.c[0] = c[0]; nd.c[1] = c[1]; nd.c[2] = c[2];
    nd.depth = depth;
    nd.index = index;
    for (int i = 0; i < 8; ++i) nd.child[i] = -1;
    pool.push_back(nd);
    return (int)pool.size() - 1;
  }
};

thread_local Tree* g_tree = nullptr;
thread_local long long g_total = 0, g_terminal = 0;

}  // namespace

extern "C" int nsvf_octree_build(const double* center, const long long* points, long long n, int depth,
                                 long long* out_total, long long* out_terminal) {
  NSVF_REQUIRE(center != nullptr && out_total != nullptr, "octree_build: null argument");
  NSVF_REQUIRE(n >= 0 && depth >= 0 && depth < 31, "octree_build: bad n=%lld or depth=%d", n, depth);
  delete g_tree;
  g_tree = new Tree();
  Tree& t = *g_tree;
  t.pool.reserve((size_t)n * 2 + 16);
  t.new_node(center, depth, -1);
  for (long long k = 0; k < n; ++k) {
    const double p[3] = {(double)points[k * 3 + 0], (double)points[k * 3 + 1], (double)points[k * 3 + 2]};
    int cur = 0;
    for (;;) {
      const int d = t.pool[cur].depth;
      const int bx = p[0] > t.pool[cur].c[0], by = p[1] > t.pool[cur].c[1], bz = p[2] > t.pool[cur].c[2];
      const int slot = bx + 2 * by + 4 * bz;
      if (d == 0) {
        const int leaf = t.new_node(p, -1, (int)k);
        t.pool[cur].child[slot] = leaf;  // replaces (and orphans) any earlier leaf in this slot
        break;
      }
      if (t.pool[cur].child[slot] < 0) {
        const double len = (double)(1 << (d - 1));
        const double nc[3] = {t.pool[cur].c[0] + (2 * bx - 1) * len, t.pool[cur].c[1] + (2 * by - 1) * len,
                              t.pool[cur].c[2] + (2 * bz - 1) * len};
        const int child = t.new_node(nc, d - 1, -1);
        t.pool[cur].child[slot] = child;
      }
      cur = t.pool[cur].child[slot];
    }
  }
  // count reachable nodes (orphaned leaves are not counted, as in EasyOctree::count)
  long long total = 0, terminal = 0;
  std::vector<int> stack{0};
  while (!stack.empty()) {
    const int cur = stack.back();
    stack.pop_back();
    ++total;
    if (t.pool[cur].depth == -1) ++terminal;
    for (int i = 0; i < 8; ++i)
      if (t.pool[cur].child[i] >= 0) stack.push_back(t.pool[cur].child[i]);
  }
  g_total = total;
  g_terminal = terminal;
  *out_total = total;
  if (out_terminal) *out_terminal = terminal;
  return 0;
}

extern "C" int nsvf_octree_flatten(int* centers, int* children, long long capacity_nodes) {
  NSVF_REQUIRE(g_tree != nullptr, "octree_flatten: call nsvf_octree_build first (same thread)");
  NSVF_REQUIRE(centers != nullptr && children != nullptr && capacity_nodes >= g_total,
               "octree_flatten: output buffers too small (%lld < %lld nodes)", capacity_nodes, g_total);
  Tree& t = *g_tree;
  const long long T = g_total;
  std::memset(centers, 0, sizeof(int) * (size_t)T * 3);
  for (long long i = 0; i < T * 9; ++i) children[i] = -1;
  long long node_idx = T - 1;
  t.pool[0].index = (int)node_idx;
  std::queue<int> q;
  q.push(0);
  while (!q.empty()) {
    const int cur = q.front();
    q.pop();
    Node& nd = t.pool[cur];
    for (int i = 0; i < 8; ++i) {
      const int ch = nd.child[i];
      if (ch < 0) continue;
      if (t.pool[ch].depth > -1) {
        --node_idx;
        t.pool[ch].index = (int)node_idx;
      }
      q.push(ch);
      if (nd.index >= 0 && nd.index < T) children[(long long)nd.index * 9 + i] = t.pool[ch].index;
    }
    if (nd.index >= 0 && nd.index < T) {
      children[(long long)nd.index * 9 + 8] = 1 << (nd.depth + 1);
      for (int a = 0; a < 3; ++a) centers[(long long)nd.index * 3 + a] = (int)nd.c[a];
    }
  }
  delete g_tree;
  g_tree = nullptr;
  return 0;
}
