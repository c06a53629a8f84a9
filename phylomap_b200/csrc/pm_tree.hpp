// Host-side tree preparation: validation of the reference's tree inputs (x$edge, nen, nodelist, root — the
// arguments of every maketreelist* entry, src/phylomap.cpp:891) and their conversion into the level schedules
// the kernels walk.  Also the O(E) replacement of the R helpers pruningwiseedgeorder / makenodelist / myreorder
// (R/sumstatMCMC.R:1-18), which are O(E^2) R loops in the reference.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace pm {
namespace host {

struct Schedule {
  int T = 0, E = 0, root = 0;              // root 0-based
  std::vector<int> e_parent, e_child;      // 0-based node ids per edge row
  std::vector<int> parent_edge;            // per node: the edge row arriving at it (-1 for the root)
  std::vector<int> up_entries, up_off;     // 5 ints per internal node (parent, a, ea, b, eb), grouped by height
  std::vector<int> down_entries, down_off; // 3 ints per drawn node (v, parent, edge), grouped by depth
};

// Throws std::string on malformed input.
inline void build_schedule(int T, int E, const int32_t* edge, const int32_t* nen, const int32_t* nodelist, int root1,
                           bool draw_tips, Schedule& s) {
  if (T < 2) throw std::string("tree needs at least 2 tips");
  if (E != 2 * T - 2) throw std::string("only binary trees are supported: nrow(edge) must be 2*length(states)-2");
  const int NN = 2 * T - 1;
  if (root1 <= T || root1 > NN) throw std::string("root must be an internal node");
  s.T = T; s.E = E; s.root = root1 - 1;
  s.e_parent.resize(E); s.e_child.resize(E);
  s.parent_edge.assign(NN, -1);
  std::vector<int> nchild(NN, 0);
  for (int e = 0; e < E; e++) {
    const int p = edge[e], c = edge[E + e];
    if (p <= T || p > NN) throw std::string("edge[,1] must hold internal nodes (T+1..2T-1)");
    if (c < 1 || c > NN) throw std::string("edge[,2] out of range");
    if (c == root1) throw std::string("the root cannot be a child");
    if (s.parent_edge[c - 1] >= 0) throw std::string("a node has two parent edges");
    s.e_parent[e] = p - 1; s.e_child[e] = c - 1;
    s.parent_edge[c - 1] = e;
    nchild[p - 1]++;
  }
  for (int v = T; v < NN; v++) if (nchild[v] != 2) throw std::string("every internal node needs exactly two children");

  // pruning order -> heights
  std::vector<int> height(NN, -1);
  for (int v = 0; v < T; v++) height[v] = 0;
  std::vector<char> used(E, 0);
  const int Nn = T - 1;
  std::vector<int> pair_parent(Nn), pair_h(Nn);
  int maxh = 0;
  for (int i = 0; i < Nn; i++) {
    const int ea = nen[2 * i] - 1, eb = nen[2 * i + 1] - 1;
    if (ea < 0 || ea >= E || eb < 0 || eb >= E || ea == eb) throw std::string("nen holds an invalid edge row");
    if (used[ea] || used[eb]) throw std::string("nen is not a permutation of the edge rows");
    used[ea] = used[eb] = 1;
    if (s.e_parent[ea] != s.e_parent[eb]) throw std::string("nen must list sibling edges next to each other");
    const int a = s.e_child[ea], b = s.e_child[eb], p = s.e_parent[ea];
    if (height[a] < 0 || height[b] < 0) throw std::string("nen is not a pruning-wise order (a child is used before it is computed)");
    height[p] = 1 + (height[a] > height[b] ? height[a] : height[b]);
    pair_parent[i] = p; pair_h[i] = height[p];
    if (height[p] > maxh) maxh = height[p];
  }
  s.up_off.assign(maxh + 1, 0);  // levels 1..maxh -> slots 0..maxh-1
  for (int i = 0; i < Nn; i++) s.up_off[pair_h[i]]++;
  {
    int acc = 0;
    for (int h = 1; h <= maxh; h++) { const int c = s.up_off[h]; s.up_off[h - 1] = acc; acc += c; }
    s.up_off[maxh] = acc;
  }
  s.up_entries.resize((size_t)5 * Nn);
  {
    std::vector<int> cur(s.up_off.begin(), s.up_off.end() - 1);
    for (int i = 0; i < Nn; i++) {
      const int ea = nen[2 * i] - 1, eb = nen[2 * i + 1] - 1;
      int* en = &s.up_entries[(size_t)5 * cur[pair_h[i] - 1]++];
      en[0] = pair_parent[i]; en[1] = s.e_child[ea]; en[2] = ea; en[3] = s.e_child[eb]; en[4] = eb;
    }
  }

  // top-down order -> depths
  std::vector<int> depth(NN, -1);
  depth[s.root] = 0;
  int maxd = 0;
  for (int i = 0; i < T - 2; i++) {
    const int v = nodelist[i] - 1;
    if (v < T || v >= NN || v == s.root) throw std::string("nodelist must hold the internal nodes below the root");
    if (depth[v] >= 0) throw std::string("nodelist repeats a node");
    const int p = s.e_parent[s.parent_edge[v]];
    if (depth[p] < 0) throw std::string("nodelist is not top-down (a node comes before its parent)");
    depth[v] = depth[p] + 1;
    if (depth[v] > maxd) maxd = depth[v];
  }
  if (draw_tips) for (int v = 0; v < T; v++) {
    depth[v] = depth[s.e_parent[s.parent_edge[v]]] + 1;
    if (depth[v] > maxd) maxd = depth[v];
  }
  std::vector<int> cnt(maxd + 1, 0);
  int ndraw = 0;
  for (int v = 0; v < NN; v++) if (depth[v] > 0) { cnt[depth[v]]++; ndraw++; }
  s.down_off.assign(maxd + 1, 0);
  {
    int acc = 0;
    for (int d = 1; d <= maxd; d++) { s.down_off[d - 1] = acc; acc += cnt[d]; }
    s.down_off[maxd] = acc;
  }
  s.down_entries.resize((size_t)3 * ndraw);
  {
    std::vector<int> cur(s.down_off.begin(), s.down_off.end() - 1);
    for (int v = 0; v < NN; v++) if (depth[v] > 0) {
      int* en = &s.down_entries[(size_t)3 * cur[depth[v] - 1]++];
      en[0] = v; en[1] = s.e_parent[s.parent_edge[v]]; en[2] = s.parent_edge[v];
    }
  }
}

// Clade schedule of the production pruning kernel (k_prune_clade).  The tree is cut into clades (complete subtrees) of
// at most `clade_max` internal nodes; each of the `nwarps` warps of a block owns a set of clades (balanced by size) and
// walks them alone in post-order, larger child subtree first, so that a parent follows its last child immediately (the
// child's partial is still in registers) and its other child shortly before (still in L2).  What lies above the clades
// ("top", ~ Nn / clade_max nodes) is processed level by level by the whole block afterwards.
//   entries: 8 ints per node: parent, a, ea, b | eb, flags, 0, 0
//   flags:   1 = child a is the node processed just before by the same warp, 2 = child b is,
//            4 = child a is internal and must be loaded, 8 = child b is
struct CladeSchedule {
  std::vector<int> entries;   // phase-1 sequences of warp 0, 1, .., then the top levels
  std::vector<int> warp_off;  // [nwarps + 1] node offsets of the warps' sequences
  std::vector<int> top_off;   // [n_top_levels + 1] node offsets of the top levels
  // the same clades top-down (k_nodes_clade): 4 ints per drawn internal node: v, parent, edge, 1 if the parent is the node
  // drawn just before by the same warp.  First the nodes above the clades by depth (the root excluded), then the warps'
  // pre-order sequences.
  std::vector<int> down_top, down_top_off;  // levels of the top part
  std::vector<int> down_seq, down_warp_off; // [nwarps + 1]
};

inline void build_clade_schedule(const Schedule& s, int nwarps, int clade_max, CladeSchedule& c) {
  const int T = s.T, NN = 2 * T - 1, Nn = T - 1;
  std::vector<int> ka(NN, -1), kb(NN, -1), ea(NN, -1), eb(NN, -1), size(NN, 0), height(NN, 0);
  std::vector<int> order(Nn);  // internal nodes, children before parents (the level order of `up_entries`)
  for (int i = 0; i < Nn; i++) {
    const int* en = &s.up_entries[(size_t)5 * i];
    const int v = en[0];
    ka[v] = en[1]; ea[v] = en[2]; kb[v] = en[3]; eb[v] = en[4];
    order[i] = v;
  }
  for (int i = 0; i < Nn; i++) {
    const int v = order[i];
    size[v] = 1 + size[ka[v]] + size[kb[v]];
  }
  if (clade_max < 1) clade_max = 1;
  // clade roots: size <= clade_max while the parent's is larger (or the node is the root)
  std::vector<int> roots;
  std::vector<char> is_top(NN, 0);
  for (int i = 0; i < Nn; i++) {
    const int v = order[i];
    if (size[v] > clade_max) { is_top[v] = 1; continue; }
    const bool at_root = v == s.root;
    const int par = at_root ? -1 : s.e_parent[s.parent_edge[v]];
    if (at_root || size[par] > clade_max) roots.push_back(v);
  }
  // longest-processing-time assignment of the clades to the warps
  std::sort(roots.begin(), roots.end(), [&](int x, int y) { return size[x] != size[y] ? size[x] > size[y] : x < y; });
  std::vector<std::vector<int>> mine(nwarps);
  std::vector<long long> load(nwarps, 0);
  for (int r : roots) {
    int w = 0;
    for (int k = 1; k < nwarps; k++) if (load[k] < load[w]) w = k;
    mine[w].push_back(r);
    load[w] += size[r];
  }
  c.entries.clear();
  c.warp_off.assign(nwarps + 1, 0);
  auto emit = [&](int v, int prev) {
    int fl = 0;
    if (ka[v] >= T) fl |= (ka[v] == prev) ? 1 : 4;
    if (kb[v] >= T) fl |= (kb[v] == prev) ? 2 : 8;
    const int en[8] = {v, ka[v], ea[v], kb[v], eb[v], fl, 0, 0};
    c.entries.insert(c.entries.end(), en, en + 8);
  };
  std::vector<int> stack;
  std::vector<char> expanded(NN, 0);
  int count = 0;
  for (int w = 0; w < nwarps; w++) {
    c.warp_off[w] = count;
    int prev = -1;
    for (int r : mine[w]) {
      stack.push_back(r);
      while (!stack.empty()) {
        const int v = stack.back();
        if (v < T) { stack.pop_back(); continue; }
        if (!expanded[v]) {
          expanded[v] = 1;
          // pushed last = processed first: the larger subtree
          const int big = size[ka[v]] >= size[kb[v]] ? ka[v] : kb[v];
          const int small = big == ka[v] ? kb[v] : ka[v];
          stack.push_back(small);
          stack.push_back(big);
        } else { stack.pop_back(); emit(v, prev); prev = v; count++; }
      }
    }
  }
  c.warp_off[nwarps] = count;
  // the top: levels by height above the clade roots; every internal child is loaded (flags 4 / 8)
  int maxh = 0;
  for (int i = 0; i < Nn; i++) {
    const int v = order[i];
    if (!is_top[v]) continue;
    const int ha = is_top[ka[v]] ? height[ka[v]] : 0, hb = is_top[kb[v]] ? height[kb[v]] : 0;
    height[v] = 1 + std::max(ha, hb);
    maxh = std::max(maxh, height[v]);
  }
  // ---- top-down (node draws) ----
  {
    std::vector<int> depth(NN, 0);
    std::vector<std::vector<int>> by_depth;
    for (int i = Nn - 1; i >= 0; i--) {  // parents before children
      const int v = order[i];
      if (!is_top[v]) continue;
      if (v != s.root) depth[v] = depth[s.e_parent[s.parent_edge[v]]] + 1;
      if (v == s.root) continue;
      if ((int)by_depth.size() <= depth[v]) by_depth.resize(depth[v] + 1);
      by_depth[depth[v]].push_back(v);
    }
    c.down_top.clear(); c.down_top_off.assign(1, 0);
    for (size_t d = 1; d < by_depth.size(); d++) {
      for (int v : by_depth[d]) {
        const int e = s.parent_edge[v];
        const int en[4] = {v, s.e_parent[e], e, 0};
        c.down_top.insert(c.down_top.end(), en, en + 4);
      }
      c.down_top_off.push_back((int)c.down_top.size() / 4);
    }
    c.down_seq.clear(); c.down_warp_off.assign(nwarps + 1, 0);
    int cnt2 = 0;
    for (int w = 0; w < nwarps; w++) {
      c.down_warp_off[w] = cnt2;
      int prev = -1;
      for (int r : mine[w]) {
        if (r == s.root) { prev = -1; }  // the root is drawn separately; its clade starts with its children
        stack.clear();
        stack.push_back(r);
        while (!stack.empty()) {
          const int v = stack.back(); stack.pop_back();
          if (v < T) continue;
          if (v != s.root) {
            const int e = s.parent_edge[v], par = s.e_parent[e];
            const int en[4] = {v, par, e, par == prev ? 1 : 0};
            c.down_seq.insert(c.down_seq.end(), en, en + 4);
            cnt2++;
            prev = v;
          }
          // pushed last = visited first: the larger subtree right after its parent
          const int big = size[ka[v]] >= size[kb[v]] ? ka[v] : kb[v];
          const int small = big == ka[v] ? kb[v] : ka[v];
          stack.push_back(small);
          stack.push_back(big);
        }
      }
    }
    c.down_warp_off[nwarps] = cnt2;
  }
  std::vector<std::vector<int>> by_height(maxh + 1);
  for (int i = 0; i < Nn; i++) if (is_top[order[i]]) by_height[height[order[i]]].push_back(order[i]);
  c.top_off.assign(maxh + 1, 0);
  for (int h = 1; h <= maxh; h++) {
    c.top_off[h - 1] = count;
    for (int v : by_height[h]) { emit(v, -1); count++; }
  }
  c.top_off[maxh] = count;
}

// A valid pruning-wise edge order (children before parents, the two edges of a node adjacent), the internal nodes
// below the root top-down (reverse pruning order, like makenodelist), and the root.  O(E), iterative.
inline void tree_order(const int32_t* edge, int E, int T, int32_t* nen, int32_t* nodelist, int32_t* root1) {
  if (T < 2 || E != 2 * T - 2) throw std::string("only binary trees are supported");
  const int NN = 2 * T - 1;
  std::vector<int> kid(2 * (size_t)NN, -1), has_parent(NN, 0);
  for (int e = 0; e < E; e++) {
    const int p = edge[e] - 1, c = edge[E + e] - 1;
    if (p < T || p >= NN || c < 0 || c >= NN) throw std::string("edge entry out of range");
    if (kid[2 * p] < 0) kid[2 * p] = e; else if (kid[2 * p + 1] < 0) kid[2 * p + 1] = e;
    else throw std::string("a node has more than two children");
    if (has_parent[c]) throw std::string("a node has two parent edges");
    has_parent[c] = 1;
  }
  int root = -1;
  for (int v = T; v < NN; v++) {
    if (kid[2 * v + 1] < 0) throw std::string("every internal node needs exactly two children");
    if (!has_parent[v]) { if (root >= 0) throw std::string("more than one root"); root = v; }
  }
  if (root < 0) throw std::string("no root");
  // iterative post-order
  std::vector<int> stack, order;  // order: internal nodes, children before parents
  std::vector<char> expanded(NN, 0);
  stack.push_back(root);
  while (!stack.empty()) {
    const int v = stack.back();
    if (v < T) { stack.pop_back(); continue; }
    if (!expanded[v]) {
      expanded[v] = 1;
      stack.push_back(edge[E + kid[2 * v + 1]] - 1);
      stack.push_back(edge[E + kid[2 * v]] - 1);
    } else { stack.pop_back(); order.push_back(v); }
  }
  if ((int)order.size() != T - 1) throw std::string("edge matrix is not a tree");
  for (int i = 0; i < T - 1; i++) { nen[2 * i] = kid[2 * order[i]] + 1; nen[2 * i + 1] = kid[2 * order[i] + 1] + 1; }
  for (int i = 1; i <= T - 2; i++) nodelist[i - 1] = order[T - 2 - i] + 1;
  *root1 = root + 1;
}

}  // namespace host
}  // namespace pm
