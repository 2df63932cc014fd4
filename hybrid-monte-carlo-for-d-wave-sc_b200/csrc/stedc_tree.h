// stedc_tree.h -- host-side description of the divide-and-conquer tree shared by all
// matrices of a batch: every node is split at every level, so all leaves sit at the
// same depth and every level is a uniform batch of merges.
#pragma once
#include <vector>

struct DcLeaf { int off, size; };
struct DcMerge { int off, n1, n2; };
struct DcTree {
  std::vector<DcLeaf> leaves;
  std::vector<std::vector<DcMerge>> levels;  // levels[0] merges pairs of leaves, last level = root
  int max_leaf = 0;
};

inline DcTree build_dc_tree(int n, int leafmax) {
  DcTree t;
  int depth = 0;
  while (((n + (1 << depth) - 1) >> depth) > leafmax) ++depth;
  std::vector<DcLeaf> cur{{0, n}};
  std::vector<std::vector<DcMerge>> top_down;
  for (int l = 0; l < depth; ++l) {
    std::vector<DcLeaf> next;
    std::vector<DcMerge> mg;
    for (auto& nd : cur) {
      int n1 = nd.size / 2, n2 = nd.size - n1;
      mg.push_back({nd.off, n1, n2});
      next.push_back({nd.off, n1});
      next.push_back({nd.off + n1, n2});
    }
    top_down.push_back(mg);
    cur = next;
  }
  t.leaves = cur;
  for (auto& lf : cur) t.max_leaf = lf.size > t.max_leaf ? lf.size : t.max_leaf;
  for (int l = depth - 1; l >= 0; --l) t.levels.push_back(top_down[l]);
  return t;
}
