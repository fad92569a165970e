// ba_nd_plan.h -- host-side plan of the PARTITIONED banded reduced solve (nested dissection of the pose chain).
// Replaces the serial column chain of `Am_BCinvBt_mat.ldlt().solve(am_BCinv_b_mat)`
// (core/full_bundle_adjustment_solver.cpp:890-908) for sequential trajectories.
//
// S is block-banded: pose j is co-visible with poses j-b .. j+b only (b = longest track span), so removing a
// SEPARATOR of b consecutive poses decouples what lies left of it from what lies right of it.  The pose chain
// is cut recursively by 2^L - 1 separators into 2^L leaves; a leaf may be cut further into a CHAIN of chunks.
// Every tree node is one dense FRONT
//
//     rows / columns = [ own (k) | right boundary Rb (wr) | left boundary Lb (wl) | rhs (1) ]
//
// whose `own` columns are eliminated by a partial Cholesky; what is left on [Rb | Lb | rhs] (the node's
// contribution block U) is added into the parent's front (multifrontal extend-add).  The boundaries of a node
// are the nearest separators to its left and right among its ancestors (for a chunk: the first b poses of the
// next chunk), both exactly where its fill can reach.  All fronts of one level are independent: the dependent
// column chain shrinks from 6N to about 6 (N / 2^L) + 6 b L columns.
//
// Pure C++ (no CUDA) so that the plan can be checked on a CPU (tests/test_nd_plan_cpu.py emulates the fronts in
// numpy against numpy.linalg.solve).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace ba {

struct NdNode {
  int own0, k, k8;      // first scalar column, number of columns, padded to 8 (padding = identity columns)
  int rb0, wr;          // right boundary: first scalar column, count (0: none)
  int lb0, wl;          // left boundary
  int b8;               // pad8(wr + wl + 1): boundary rows incl. the rhs row (front-local index k8 + wr + wl)
  int child[2];         // -1: none.  child[0]: chain predecessor or left subtree, child[1]: right subtree
  int parent;
  int rb_off, lb_off, rhs_off;   // this node's U rows in the PARENT's front-local index space
  int level;            // forward level (children have smaller levels); backward runs the levels in reverse
  int cta, seq;         // persistent kernel: CTA that owns the node and position in that CTA's list
  int helper;           // persistent driver: CTA that holds this front's boundary x boundary accumulators (-1: the
                        // front's own CTA does); the main CTA publishes one flag per panel step (step_flag0 + s)
  int step_flag0;
  int bandT;            // row tiles below the diagonal tile of a column that can be nonzero in [own | Rb] (chain fronts:
                        // band of S; separators: all).  Rows outside (except Lb / rhs) are skipped by the factorisation
  long long L_off;      // doubles: factor tiles of the own columns [n_tiles][64] (swizzled 8 x 8 tiles, column by
                        // column), then the inverses of the diagonal blocks [KT][64]
  long long U_off;      // doubles: contribution block IN THE PARENT'S FRONT LAYOUT: the parent's own-column tiles
                        // [n_tiles_parent][64] (swizzled like the shared-memory window) followed by the parent's
                        // boundary x boundary tiles [BT_p (BT_p + 1) / 2][64] (plain).  The child scatters its
                        // Schur complement there, the parent's assembly is then a streaming add of whole tiles
};

struct NdPlan {
  bool valid = false;
  int n = 0, b = 0, bw = 0;     // matrix order, separator width in poses, scalar half-bandwidth
  int depth = 0, n_leaves = 0, n_levels = 0, n_ctas = 0;
  int max_KT = 0, max_BT = 0, max_R8 = 0, max_tiles = 0, max_list = 0;
  size_t smem_bytes = 0;
  long long L_doubles = 0, U_doubles = 0;
  std::vector<NdNode> nodes;
  std::vector<int> level_ptr, level_nodes;   // nodes by level
  std::vector<int> cta_ptr, cta_nodes;       // persistent kernel: per CTA, its nodes in forward order
  int n_helpers = 0, n_step_flags = 0;       // helper CTAs (blockIdx >= n_ctas) and their per-step flags
  std::vector<int> helper_nodes;             // front of every helper CTA
};

inline int nd_pad8(int v) { return (v + 7) / 8 * 8; }
constexpr int kNdMaxBT = 22;            // boundary tile rows a consumer warp can hold in registers (9 warps x 30 tiles)
constexpr size_t kNdSmemLimit = 227 * 1024;
constexpr int kNdHelperMinBT = 12;      // boundary tile rows from which a front gets a helper CTA
constexpr int kNdMaxSteps = 32;         // per-step flags reserved per helped front

// shared memory of one front: own trapezoid tiles + inverses of the diagonal blocks + x / rhs vectors of the backward
// pass + child index maps + tile table (+ node records)
inline size_t nd_front_smem(int KT, int BT) {
  const size_t tiles = (size_t)KT * (KT + 1) / 2 + (size_t)BT * KT;
  const size_t R8 = 8 * (size_t)(KT + BT);
  return (tiles + KT) * 64 * sizeof(double) + 2 * R8 * sizeof(double) + 2 * R8 * sizeof(int) + tiles * sizeof(int) +
         1024;
}

// N free poses (n = 6N), b = largest pose distance inside a track (scalar half-bandwidth 6b + 5).
// max_ctas: CTAs that can be co-resident (one per SM).  Returns pl.valid = false when the band is too wide or the
// chain too short for a partition to pay.
constexpr int kNdTotalCtas = 144;       // persistent launch: CTAs that must be co-resident (one per SM, B200 has 148)
inline void nd_make_plan_depth(NdPlan &pl, int N, int b, int max_ctas, int force_depth, int force_chunk) {
  pl = NdPlan();
  pl.n = 6 * N; pl.b = b; pl.bw = 6 * b + 5;
  if (N <= 0 || b <= 0 || 6 * N >= 65000) return;   // the kernels address S with 32-bit element offsets
  const int w = 6 * b;
  const int BT = nd_pad8(2 * w + 1) / 8;
  if (BT > kNdMaxBT) return;
  const int sepKT = nd_pad8(w) / 8;
  if (nd_front_smem(sepKT, BT) > kNdSmemLimit) return;
  // largest chunk (poses) whose front fits in shared memory
  int max_chunk = b;
  while (nd_front_smem(nd_pad8(6 * (max_chunk + 1)) / 8, BT) <= kNdSmemLimit) ++max_chunk;
  if (force_chunk > 0) max_chunk = std::max(b, std::min(max_chunk, force_chunk));
  int best_depth = -1;
  best_depth = force_depth;
  if (best_depth < 1) return;
  {  // feasibility: every leaf holds at least b poses (separators on both sides of a leaf must not couple)
    const long long P = 1LL << best_depth;
    if (P > max_ctas || (long long)N - (P - 1) * b < P * (long long)std::max(b, 1)) return;
  }
  pl.depth = best_depth;
  pl.n_leaves = 1 << best_depth;

  struct Range { int lo, hi; };   // poses
  bool chunk_overflow = false;
  std::vector<NdNode> &nodes = pl.nodes;
  auto new_node = [&](int p0, int p1, Range lb, Range rb) {
    NdNode nd{};
    nd.own0 = 6 * p0; nd.k = 6 * (p1 - p0); nd.k8 = nd_pad8(nd.k);
    nd.rb0 = 6 * rb.lo; nd.wr = 6 * (rb.hi - rb.lo);
    nd.lb0 = 6 * lb.lo; nd.wl = 6 * (lb.hi - lb.lo);
    nd.b8 = nd_pad8(nd.wr + nd.wl + 1);
    nd.child[0] = nd.child[1] = -1;
    nd.parent = -1;
    nd.rb_off = nd.lb_off = nd.rhs_off = -1;
    nodes.push_back(nd);
    return (int)nodes.size() - 1;
  };
  auto link = [&](int c, int p, int slot, int rb_off, int lb_off) {
    nodes[c].parent = p;
    nodes[p].child[slot] = c;
    nodes[c].rb_off = rb_off;
    nodes[c].lb_off = lb_off;
    nodes[c].rhs_off = nodes[p].k8 + nodes[p].wr + nodes[p].wl;
  };
  // leaf [lo, hi) as a chain of chunks; returns the LAST chunk (the one whose Rb is the separator rb)
  auto build_leaf = [&](int lo, int hi, Range lb, Range rb) {
    const int len = hi - lo;
    int nchunk = std::max(1, (len + max_chunk - 1) / max_chunk);
    while (nchunk > 1 && len / nchunk < b) --nchunk;   // every chunk holds the previous chunk's right boundary
    if ((len + nchunk - 1) / nchunk > max_chunk) chunk_overflow = true;
    int prev = -1, p0 = lo;
    for (int c = 0; c < nchunk; ++c) {
      const int p1 = (c == nchunk - 1) ? hi : lo + (int)((long long)len * (c + 1) / nchunk);
      const Range r = (c == nchunk - 1) ? rb : Range{p1, std::min(p1 + b, hi)};
      const int id = new_node(p0, p1, lb, r);
      if (prev >= 0) link(prev, id, 0, 0, nodes[id].k8 + nodes[id].wr);
      prev = id;
      p0 = p1;
    }
    return prev;
  };
  // recursive bisection
  struct Rec {
    std::vector<NdNode> &nodes;
    int b;
    decltype(new_node) &mk;
    decltype(link) &lk;
    decltype(build_leaf) &leaf;
    int build(int lo, int hi, Range lb, Range rb, int d) {
      if (d == 0) return leaf(lo, hi, lb, rb);
      const int len = hi - lo - b;
      const int s0 = lo + len / 2;
      const Range sep{s0, s0 + b};
      const int l = build(lo, s0, lb, sep, d - 1);
      const int r = build(s0 + b, hi, sep, rb, d - 1);
      const int id = mk(s0, s0 + b, lb, rb);
      lk(l, id, 0, 0, nodes[id].k8 + nodes[id].wr);   // left: its Rb is my own block, its Lb is my Lb
      lk(r, id, 1, nodes[id].k8, 0);                  // right: its Lb is my own block, its Rb is my Rb
      return id;
    }
  } rec{nodes, b, new_node, link, build_leaf};
  const int root = rec.build(0, N, Range{0, 0}, Range{0, 0}, best_depth);
  (void)root;

  // levels (children first), workspace offsets, per-CTA lists
  const int nn = (int)nodes.size();
  std::vector<int> order(nn);
  for (int i = 0; i < nn; ++i) {   // children are created before their parents
    int lv = 0;
    for (int c = 0; c < 2; ++c)
      if (nodes[i].child[c] >= 0) lv = std::max(lv, nodes[nodes[i].child[c]].level + 1);
    nodes[i].level = lv;
    pl.n_levels = std::max(pl.n_levels, lv + 1);
  }
  pl.level_ptr.assign(pl.n_levels + 1, 0);
  for (int i = 0; i < nn; ++i) pl.level_ptr[nodes[i].level + 1]++;
  for (int l = 0; l < pl.n_levels; ++l) pl.level_ptr[l + 1] += pl.level_ptr[l];
  pl.level_nodes.assign(nn, 0);
  {
    std::vector<int> fill(pl.level_ptr.begin(), pl.level_ptr.end() - 1);
    for (int i = 0; i < nn; ++i) pl.level_nodes[fill[nodes[i].level]++] = i;
  }
  for (int i = 0; i < nn; ++i) {
    NdNode &nd = nodes[i];
    const int KT = nd.k8 / 8, BTn = nd.b8 / 8, R8 = nd.k8 + nd.b8;
    nd.L_off = pl.L_doubles;
    pl.L_doubles += ((long long)KT * (KT + 1) / 2 + (long long)BTn * KT + KT) * 64;
    nd.U_off = pl.U_doubles;
    if (nd.parent >= 0) {
      const NdNode &pn = nodes[nd.parent];
      const long long KTp = pn.k8 / 8, BTp = pn.b8 / 8;
      pl.U_doubles += (KTp * (KTp + 1) / 2 + BTp * KTp + BTp * (BTp + 1) / 2) * 64;
    }
    pl.max_KT = std::max(pl.max_KT, KT);
    pl.max_BT = std::max(pl.max_BT, BTn);
    pl.max_R8 = std::max(pl.max_R8, R8);
    pl.max_tiles = std::max(pl.max_tiles, KT * (KT + 1) / 2 + BTn * KT);
    pl.smem_bytes = std::max(pl.smem_bytes, nd_front_smem(KT, BTn));
  }
  for (int i = 0; i < nn; ++i) {
    NdNode &nd = nodes[i];
    nd.bandT = (nd.child[1] >= 0) ? (1 << 20) : (7 + pl.bw + (nd.k8 - nd.k) + 7) / 8;
  }
  // CTA lists.  A CTA starts at the first chunk of a leaf and climbs while it arrives through child[0]; fronts near
  // the root get CTAs of their own while the launch stays co-resident (all of them when the tree is small): their
  // factor is then still in shared memory when the backward pass comes down the tree.
  {
    std::vector<char> own(nn, 0);
    int n_first = 0;
    for (int i = 0; i < nn; ++i)
      if (nodes[i].child[0] < 0 && nodes[i].child[1] < 0) ++n_first;
    int budget = std::max(0, std::max(kNdTotalCtas, max_ctas) - n_first);
    for (int l = pl.n_levels - 1; l >= 1 && budget > 0; --l)
      for (int q = pl.level_ptr[l]; q < pl.level_ptr[l + 1] && budget > 0; ++q) {
        own[pl.level_nodes[q]] = 1;
        --budget;
      }
    int n_cta = 0;
    for (int i = 0; i < nn; ++i) nodes[i].cta = -1;
    pl.cta_ptr.assign(1, 0);
    for (int i = 0; i < nn; ++i) {
      const bool first = nodes[i].child[0] < 0 && nodes[i].child[1] < 0;
      if (!first && !own[i]) continue;
      int t = i, seq = 0;
      for (;;) {
        nodes[t].cta = n_cta;
        nodes[t].seq = seq++;
        pl.cta_nodes.push_back(t);
        const int p = nodes[t].parent;
        if (p < 0 || nodes[p].child[0] != t || own[p]) break;
        t = p;
      }
      pl.cta_ptr.push_back((int)pl.cta_nodes.size());
      pl.max_list = std::max(pl.max_list, seq);
      ++n_cta;
    }
    pl.n_ctas = n_cta;
    // Helper CTAs: on a large front half of the step is the rank-8 update of the boundary x boundary block, which is
    // off the critical path (only the parent needs it).  A second CTA on another SM holds those accumulators, follows
    // the main CTA panel by panel through per-step flags and reads the solved boundary tiles from the factor in global
    // memory; the main CTA is left with the panel solve and the own-column update.
    for (int i = 0; i < nn; ++i) { nodes[i].helper = -1; nodes[i].step_flag0 = -1; }
    std::vector<int> cand;
    for (int i = 0; i < nn; ++i)
      if (nodes[i].b8 / 8 >= kNdHelperMinBT && nodes[i].k8 / 8 >= 4 && nodes[i].k8 / 8 <= kNdMaxSteps && nodes[i].parent >= 0) cand.push_back(i);
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
      return (long long)nodes[a].b8 * nodes[a].b8 * nodes[a].k8 > (long long)nodes[b].b8 * nodes[b].b8 * nodes[b].k8; });
    for (int i : cand) {
      if (n_cta + pl.n_helpers >= std::max(kNdTotalCtas, max_ctas)) break;
      nodes[i].helper = n_cta + pl.n_helpers;
      nodes[i].step_flag0 = pl.n_step_flags;
      pl.n_step_flags += kNdMaxSteps;
      pl.helper_nodes.push_back(i);
      ++pl.n_helpers;
    }
  }
  if (pl.smem_bytes > kNdSmemLimit || chunk_overflow) return;
  pl.valid = true;
}

// estimated length of the dependent chain (8-column panel steps + hand-overs) of a valid plan
inline double nd_plan_cost(const NdPlan &pl) {
  std::vector<double> at(pl.nodes.size(), 0.0);
  double worst = 0.0;
  for (size_t i = 0; i < pl.nodes.size(); ++i) {   // children precede parents
    double c = 0.0;
    for (int q = 0; q < 2; ++q)
      if (pl.nodes[i].child[q] >= 0) c = std::max(c, at[pl.nodes[i].child[q]]);
    at[i] = c + pl.nodes[i].k8 / 8.0 + 3.0;
    worst = std::max(worst, at[i]);
  }
  return worst;
}

inline void nd_make_plan(NdPlan &pl, int N, int b, int max_ctas = 128, int force_depth = -1, int force_chunk = -1) {
  if (force_depth >= 0) {
    nd_make_plan_depth(pl, N, b, max_ctas, force_depth, force_chunk);
    return;
  }
  pl = NdPlan();
  double best = 1e300;
  for (int d = 1; d <= 12 && (1 << d) <= max_ctas; ++d) {
    NdPlan cand;
    nd_make_plan_depth(cand, N, b, max_ctas, d, force_chunk);
    if (!cand.valid) continue;
    const double c = nd_plan_cost(cand);
    if (c < best) { best = c; pl = std::move(cand); }
  }
  // the serial window kernel walks 6N / 8 panel steps: the partition must at least halve that
  if (pl.valid && best > 0.5 * (6.0 * N / 8.0)) pl = NdPlan();
}

}  // namespace ba
