// ba_cholesky_nd.cuh -- K5 for banded reduced camera systems, PARTITIONED: nested dissection of the pose chain
// into independent dense fronts (plan: ba_nd_plan.h), one CTA per front, FP64 tensor-core (DMMA) trailing updates.
// Replaces `Am_BCinvBt_mat.ldlt().solve(am_BCinv_b_mat)` (core/full_bundle_adjustment_solver.cpp:890-908).
//
// The serial banded kernel (ba_cholesky_banded.cuh) walks n/8 dependent panel steps on ONE SM.  Here the pose
// chain is cut by separators into 2^L leaves that are eliminated concurrently on 2^L SMs; the separators are
// eliminated level by level (L levels).  The dependent chain is 6 (N / 2^L) / 8 + L (6 b / 8) panel steps.
//
// One front = [own | Rb | Lb | rhs] (ba_nd_plan.h).  Per front, one CTA of 12 warps:
//   * assembly: the own columns (own x own lower trapezoid and boundary x own) are gathered into shared memory as
//     8 x 8 tiles: entries of S (global, column-major lower + rhs row) plus the contribution blocks U of the
//     children (extend-add through per-child index maps); the boundary x boundary part never enters shared
//     memory: it lives in the DMMA accumulator registers of the consumer warps for the whole front.
//   * factor: panel steps of 8 columns.  Warp 3 (alone on its SM sub-partition) factors the 8 x 8 diagonal block
//     redundantly in every lane and applies the triangular solve to the rows below (its lanes own rows); the nine
//     consumer warps apply the rank-8 update with two DMMAs per tile: the tiles of the next panel first (then the
//     panel warp may go on), then the rest of the trapezoid, then their boundary x boundary tiles in registers.
//     Tiles have static owners, so no CTA-wide barrier is needed between steps.
//   * the panel (L11, L21 and the forward-substituted rhs row z) goes to global memory for the backward pass, the
//     accumulators are written out as the contribution block U.
// Backward (root first): x_own = -L11^-T (L21^T x_boundary - z): boundary part in parallel over all warps,
// then an 8-column block chain inside one warp with explicit inverses of the diagonal blocks.
//
// Two drivers over the same per-front code: one launch per tree level (`k_nd_forward_level` / `k_nd_backward_level`)
// and a single persistent launch (`k_nd_persistent`) in which CTA p owns leaf p, climbs the tree while it arrives as
// a left child, and hands over through acquire / release flags in global memory.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "ba_cholesky_banded.cuh"   // dmma_884b, fast_rcp, fast_rsqrt
#include "ba_nd_plan.h"

namespace ba {

constexpr int kNdThreads = 384;
constexpr int kNdCons = 9;                           // consumer warps 0,1,2,4,5,6,8,9,10
constexpr int kNdActive = 32 * (kNdCons + 1);        // consumers + panel warp (named barrier population)
constexpr int kNdSpinLimit = 1 << 22;

struct NdArgs {
  const NdNode *nodes;
  const int *list;        // level driver: node ids by level; persistent driver: per-CTA lists
  const int *list_ptr;    // persistent driver: [n_ctas + 1]
  const double *S;        // (n+1)^2, entry (r, c), r >= c, at S[c * ld + r]; rhs in row n
  int n, ld, bw;
  double *Lws, *Uws, *x;
  int *flags;             // persistent driver: [n_nodes] forward done, [n_nodes] backward done, abort, sticky error
  int n_nodes;
  int max_R8, max_tiles;  // shared-memory carve-up
};

__device__ unsigned long long g_nd_dbg[16];

__device__ __forceinline__ int nd_tidx(int I, int J) { return I * (I + 1) / 2 + J; }   // lower-triangular tile index

struct NdSmem {
  double *win;    // own trapezoid tiles [max_tiles][64]
  double *Lp;     // published panels [2][max_R8][12]
  double *xs;     // backward: x by front-local index [max_R8]
  int *pm;        // child index maps [2][max_R8]
  NdNode *node;   // current node
  NdNode *cn;     // its two children [2]
};

__device__ __forceinline__ NdSmem nd_carve(unsigned char *raw, const NdArgs &g) {
  NdSmem sm;
  sm.win = reinterpret_cast<double *>(raw);
  sm.Lp = sm.win + (size_t)g.max_tiles * 64;
  sm.xs = sm.Lp + (size_t)2 * g.max_R8 * kNdLpStride;
  sm.pm = reinterpret_cast<int *>(sm.xs + g.max_R8);
  sm.node = reinterpret_cast<NdNode *>(sm.pm + 2 * g.max_R8);
  sm.cn = sm.node + 1;
  return sm;
}
inline size_t nd_smem_bytes(const NdPlan &pl) {
  return (size_t)pl.max_tiles * 64 * sizeof(double) + (size_t)2 * pl.max_R8 * kNdLpStride * sizeof(double) +
         (size_t)pl.max_R8 * sizeof(double) + (size_t)2 * pl.max_R8 * sizeof(int) + 3 * sizeof(NdNode) + 64;
}

// element (a, b), a >= b, of a node's contribution block
__device__ __forceinline__ double nd_load_U(const double *U, int a, int b) {
  return __ldcg(U + (size_t)nd_tidx(a >> 3, b >> 3) * 64 + (a & 7) * 8 + (b & 7));
}

// initial value of front entry (i, j), i >= j (front-local indices): S part (own columns only) + children
__device__ __forceinline__ double nd_front_value(const NdArgs &g, const NdSmem &sm, int i, int j) {
  const NdNode &nd = *sm.node;
  double v = 0.0;
  if (j < nd.k8) {
    if (j >= nd.k) return (i == j) ? 1.0 : 0.0;       // identity padding of the own block
    if (i >= nd.k && i < nd.k8) return 0.0;
    const int gj = nd.own0 + j;
    const int bi = i - nd.k8;
    int gi = -1;
    if (i < nd.k) gi = nd.own0 + i;
    else if (bi < nd.wr) gi = nd.rb0 + bi;
    else if (bi < nd.wr + nd.wl) gi = nd.lb0 + (bi - nd.wr);
    else if (bi == nd.wr + nd.wl) gi = g.n;
    if (gi < 0) return 0.0;
    if (gi == g.n) {
      v = __ldcg(g.S + (size_t)gj * g.ld + g.n);
    } else {
      const int lo = min(gi, gj), hi = max(gi, gj);
      if (hi - lo <= g.bw) v = __ldcg(g.S + (size_t)lo * g.ld + hi);
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (nd.child[c] < 0) continue;
    const int pi = sm.pm[c * g.max_R8 + i], pj = sm.pm[c * g.max_R8 + j];
    if (pi >= 0 && pj >= 0) v += nd_load_U(g.Uws + sm.cn[c].U_off, max(pi, pj), min(pi, pj));
  }
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// forward elimination of one front.  TPW: boundary x boundary tiles per consumer warp (register accumulators)
// ------------------------------------------------------------------------------------------------------------
template <int TPW>
__device__ void nd_forward_node(const NdArgs &g, const NdSmem &sm, int node_id, int timing) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool is_producer = warp == 3;
  const bool is_idle = (warp & 3) == 3 && !is_producer;
  const int cwi = warp - (warp >> 2);            // consumer index 0..8
  const int ct = cwi * 32 + lane;                // consumer thread 0..287
  if (t < (int)(sizeof(NdNode) / sizeof(int))) reinterpret_cast<int *>(sm.node)[t] = reinterpret_cast<const int *>(g.nodes + node_id)[t];
  __syncthreads();
  const NdNode &nd = *sm.node;
  for (int c = 0; c < 2; ++c) {
    if (nd.child[c] >= 0 && t < (int)(sizeof(NdNode) / sizeof(int)))
      reinterpret_cast<int *>(sm.cn + c)[t] = reinterpret_cast<const int *>(g.nodes + nd.child[c])[t];
  }
  const int KT = nd.k8 >> 3, BT = nd.b8 >> 3, NT = KT + BT, R8 = 8 * NT;
  for (int i = t; i < 2 * g.max_R8; i += kNdThreads) sm.pm[i] = -1;
  __syncthreads();
  for (int c = 0; c < 2; ++c) {
    if (nd.child[c] < 0) continue;
    const NdNode &cn = sm.cn[c];
    int *pm = sm.pm + c * g.max_R8;
    for (int i = t; i < cn.wr; i += kNdThreads) pm[cn.rb_off + i] = i;
    for (int i = t; i < cn.wl; i += kNdThreads) pm[cn.lb_off + i] = cn.wr + i;
    if (t == 0) pm[cn.rhs_off] = cn.wr + cn.wl;
  }
  __syncthreads();
  auto colbase = [&](int J) { return J * NT - (J * (J - 1)) / 2; };   // tile (I, J), I >= J, at colbase(J) + I - J
  // ---- assembly of the own columns (coalesced along the rows of a column of S)
  for (int J = 0; J < KT; ++J) {
    const int nrows = R8 - 8 * J;
    double *colt = sm.win + (size_t)colbase(J) * 64;
    for (int e = t; e < 8 * nrows; e += kNdThreads) {
      const int jj = e / nrows, ri = e - jj * nrows;
      const int i = 8 * J + ri, j = 8 * J + jj;
      const double v = (i >= j) ? nd_front_value(g, sm, i, j) : 0.0;
      colt[(size_t)(ri >> 3) * 64 + (ri & 7) * 8 + jj] = v;
    }
  }
  // ---- boundary x boundary accumulators (consumers): children's contributions passed through
  const int fr = lane >> 2, fc = 2 * (lane & 3), kq = lane & 3;
  const int n_utiles = BT * (BT + 1) / 2;
  double2 acc[TPW];
  int ubi[TPW], ubj[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    acc[q] = make_double2(0.0, 0.0);
    ubi[q] = 0; ubj[q] = 0;
  }
  if (!is_producer && !is_idle) {
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      int e = cwi + kNdCons * q;
      if (e < n_utiles) {
        int I = 0;
        while (e > I) { e -= I + 1; ++I; }     // row I holds I + 1 tiles
        ubi[q] = I; ubj[q] = e;
        if (nd.child[0] >= 0 || nd.child[1] >= 0) {
          const int i = nd.k8 + 8 * I + fr, j = nd.k8 + 8 * e + fc;
          if (i >= j) acc[q].x = nd_front_value(g, sm, i, j);
          if (i >= j + 1) acc[q].y = nd_front_value(g, sm, i, j + 1);
        }
      } else {
        ubi[q] = -1;
      }
    }
  }
  __syncthreads();
  unsigned long long t0 = 0;
  if (timing && t == 0) t0 = gtime();

  // ---- factorisation of the own columns
  double *Lg = g.Lws + nd.L_off;
  const int LpBuf = g.max_R8 * kNdLpStride;
  if (is_idle) {
    // nothing
  } else if (is_producer) {
#pragma unroll 1
    for (int s = 0; s < KT; ++s) {
      const int par = s & 1;
      if (s > 0) asm volatile("bar.sync %0, %1;" ::"r"(5 + par), "n"(kNdActive) : "memory");
      const int nb = R8 - 8 * (s + 1);                    // rows below the diagonal block
      const double *col = sm.win + (size_t)colbase(s) * 64;   // tile (s + q, s) at col + 64 q
      double *LpS = sm.Lp + par * LpBuf;
      double a[4][8];
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int pos = lane + 32 * sl;
        if (pos < nb) {
          const double4 *src = reinterpret_cast<const double4 *>(col + (size_t)(1 + (pos >> 3)) * 64 + (pos & 7) * 8);
          const double4 v0 = src[0], v1 = src[1];
          a[sl][0] = v0.x; a[sl][1] = v0.y; a[sl][2] = v0.z; a[sl][3] = v0.w;
          a[sl][4] = v1.x; a[sl][5] = v1.y; a[sl][6] = v1.z; a[sl][7] = v1.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) a[sl][k] = 0.0;
        }
      }
      double D[8][8], rc[8], rs[8];
      {
        const double4 *dsrc = reinterpret_cast<const double4 *>(col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double4 v0 = dsrc[2 * i], v1 = dsrc[2 * i + 1];
          D[i][0] = v0.x; D[i][1] = v0.y; D[i][2] = v0.z; D[i][3] = v0.w;
          D[i][4] = v1.x; D[i][5] = v1.y; D[i][6] = v1.z; D[i][7] = v1.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double d = D[k][k];
        const bool pos_def = d > 0.0;
        rc[k] = pos_def ? fast_rcp(d) : 0.0;          // non-positive pivot: LDLT's D^+ = 0
        rs[k] = pos_def ? fast_rsqrt(d) : 0.0;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) {
          const double ti = D[i][k] * rc[k];
#pragma unroll
          for (int j = k + 1; j <= i; ++j) D[i][j] -= ti * D[j][k];
        }
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          const double u = a[sl][k] * rc[k];
#pragma unroll
          for (int m = k + 1; m < 8; ++m) a[sl][m] -= u * D[m][k];
        }
      }
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int pos = lane + 32 * sl;
        if (pos < nb) {
          double4 *dst = reinterpret_cast<double4 *>(LpS + (size_t)(8 * (s + 1) + pos) * kNdLpStride);
          dst[0] = make_double4(a[sl][0] * rs[0], a[sl][1] * rs[1], a[sl][2] * rs[2], a[sl][3] * rs[3]);
          dst[1] = make_double4(a[sl][4] * rs[4], a[sl][5] * rs[5], a[sl][6] * rs[6], a[sl][7] * rs[7]);
        }
      }
      // the diagonal block's own rows: L_ss (lower), for the global copy
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (lane == i) {
          double l[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) l[m] = (m < i) ? D[i][m] * rs[m] : (m == i ? D[i][i] * rs[i] : 0.0);
          double4 *dst = reinterpret_cast<double4 *>(LpS + (size_t)(8 * s + i) * kNdLpStride);
          dst[0] = make_double4(l[0], l[1], l[2], l[3]);
          dst[1] = make_double4(l[4], l[5], l[6], l[7]);
        }
      }
      // rows beyond the first 128: same multipliers, two slices at a time
#pragma unroll 1
      for (int p0 = 128; p0 < nb; p0 += 64) {
        double b[2][8];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int pos = p0 + lane + 32 * sl;
          if (pos < nb) {
            const double4 *src = reinterpret_cast<const double4 *>(col + (size_t)(1 + (pos >> 3)) * 64 + (pos & 7) * 8);
            const double4 v0 = src[0], v1 = src[1];
            b[sl][0] = v0.x; b[sl][1] = v0.y; b[sl][2] = v0.z; b[sl][3] = v0.w;
            b[sl][4] = v1.x; b[sl][5] = v1.y; b[sl][6] = v1.z; b[sl][7] = v1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) b[sl][k] = 0.0;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const double u = b[sl][k] * rc[k];
#pragma unroll
            for (int m = k + 1; m < 8; ++m) b[sl][m] -= u * D[m][k];
          }
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int pos = p0 + lane + 32 * sl;
          if (pos < nb) {
            double4 *dst = reinterpret_cast<double4 *>(LpS + (size_t)(8 * (s + 1) + pos) * kNdLpStride);
            dst[0] = make_double4(b[sl][0] * rs[0], b[sl][1] * rs[1], b[sl][2] * rs[2], b[sl][3] * rs[3]);
            dst[1] = make_double4(b[sl][4] * rs[4], b[sl][5] * rs[5], b[sl][6] * rs[6], b[sl][7] * rs[7]);
          }
        }
      }
      asm volatile("bar.arrive %0, %1;" ::"r"(3 + par), "n"(kNdActive) : "memory");
    }
  } else {
    // ---------------------------------------------- consumers ----------------------------------------------
    const int laneL = fr * kNdLpStride + kq, laneC = fr * 8 + fc;
    auto column_update = [&](const double *LpS, int J) {
      // tiles (I, J) owned by this warp: (I + 4 J) mod 9 == cwi
      int I = J + ((cwi - 5 * J) % kNdCons + kNdCons) % kNdCons;
      const double bf0 = LpS[(size_t)J * 8 * kNdLpStride + laneL], bf1 = LpS[(size_t)J * 8 * kNdLpStride + laneL + 4];
      double *colt = sm.win + (size_t)(colbase(J) - J) * 64 + laneC;   // tile (I, J) at colt + 64 I
      for (; I < NT; I += 4 * kNdCons) {
        double2 cv[4];
        double af[4][2];
        bool live[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int Iq = I + kNdCons * q;
          live[q] = Iq < NT;
          const int Ic = live[q] ? Iq : I;
          cv[q] = *reinterpret_cast<const double2 *>(colt + (size_t)Ic * 64);
          af[q][0] = -LpS[(size_t)Ic * 8 * kNdLpStride + laneL];
          af[q][1] = -LpS[(size_t)Ic * 8 * kNdLpStride + laneL + 4];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma_884b(cv[q].x, cv[q].y, af[q][0], bf0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma_884b(cv[q].x, cv[q].y, af[q][1], bf1);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (live[q]) *reinterpret_cast<double2 *>(colt + (size_t)(I + kNdCons * q) * 64) = cv[q];
      }
    };
#pragma unroll 1
    for (int s = 0; s < KT; ++s) {
      const int par = s & 1;
      const double *LpS = sm.Lp + par * LpBuf;
      asm volatile("bar.sync %0, %1;" ::"r"(3 + par), "n"(kNdActive) : "memory");
      if (s + 1 < KT) {
        column_update(LpS, s + 1);                       // the next panel first
        asm volatile("bar.arrive %0, %1;" ::"r"(5 + (par ^ 1)), "n"(kNdActive) : "memory");
      }
      for (int J = s + 2; J < KT; ++J) column_update(LpS, J);
      // boundary x boundary tiles (registers)
      const double *LpB = LpS + (size_t)KT * 8 * kNdLpStride + laneL;
#pragma unroll
      for (int q = 0; q < TPW; ++q) {
        if (ubi[q] >= 0) {
          const double *pa = LpB + (size_t)ubi[q] * 8 * kNdLpStride, *pb = LpB + (size_t)ubj[q] * 8 * kNdLpStride;
          const double a0 = -pa[0], a1 = -pa[4], b0 = pb[0], b1 = pb[4];
          dmma_884b(acc[q].x, acc[q].y, a0, b0);
          dmma_884b(acc[q].x, acc[q].y, a1, b1);
        }
      }
      // finished panel (rows 8 s .. R8 - 1) to global memory
      for (int r = 8 * s + ct; r < R8; r += 32 * kNdCons) {
        const double4 *src = reinterpret_cast<const double4 *>(LpS + (size_t)r * kNdLpStride);
        double4 *dst = reinterpret_cast<double4 *>(Lg + ((size_t)s * R8 + r) * 8);
        dst[0] = src[0];
        dst[1] = src[1];
      }
    }
    // ---- contribution block
    double *Ug = g.Uws + nd.U_off;
#pragma unroll
    for (int q = 0; q < TPW; ++q)
      if (ubi[q] >= 0) *reinterpret_cast<double2 *>(Ug + (size_t)nd_tidx(ubi[q], ubj[q]) * 64 + laneC) = acc[q];
  }
  __syncthreads();
  if (timing && t == 0) {
    atomicAdd(&g_nd_dbg[2 + min(nd.level, 6)], gtime() - t0);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward substitution of one front: x_own from x_boundary
// ------------------------------------------------------------------------------------------------------------
__device__ void nd_backward_node(const NdArgs &g, const NdSmem &sm, int node_id) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t < (int)(sizeof(NdNode) / sizeof(int))) reinterpret_cast<int *>(sm.node)[t] = reinterpret_cast<const int *>(g.nodes + node_id)[t];
  __syncthreads();
  const NdNode &nd = *sm.node;
  const int KT = nd.k8 >> 3, BT = nd.b8 >> 3, R8 = 8 * (KT + BT), k8 = nd.k8;
  const double *Lg = g.Lws + nd.L_off;
  double *Ls = sm.win;                       // own x own part of L: panel s rows [8 s, k8) at ls_off(s)
  double *Li = sm.Lp;                        // inverses of the diagonal blocks [KT][8][8]
  double *tb = sm.Lp + (size_t)KT * 64;      // right-hand side of the block chain [k8]
  auto ls_off = [&](int s) { return 8 * (s * k8 - 4 * s * (s - 1)); };
  for (int i = t; i < R8; i += kNdThreads) {
    double v = 0.0;
    const int bi = i - k8;
    if (bi >= 0) {
      if (bi < nd.wr) v = __ldcg(g.x + nd.rb0 + bi);
      else if (bi < nd.wr + nd.wl) v = __ldcg(g.x + nd.lb0 + (bi - nd.wr));
      else if (bi == nd.wr + nd.wl) v = -1.0;
    }
    sm.xs[i] = v;
  }
  for (int s = 0; s < KT; ++s) {
    const int cnt = (k8 - 8 * s) * 8;
    const double *src = Lg + ((size_t)s * R8 + 8 * s) * 8;
    double *dst = Ls + ls_off(s);
    for (int e = t; e < cnt; e += kNdThreads) dst[e] = __ldcg(src + e);
  }
  __syncthreads();
  // boundary part: tb[c] = sum_{r >= k8} L[r][c] xs[r]   (warp per panel, lanes over rows)
  for (int s = warp; s < KT; s += kNdThreads / 32) {
    double p[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = 0.0;
    for (int r = k8 + lane; r < R8; r += 32) {
      const double xv = sm.xs[r];
      const double2 *src = reinterpret_cast<const double2 *>(Lg + ((size_t)s * R8 + r) * 8);
      const double2 w0 = __ldcg(src), w1 = __ldcg(src + 1), w2 = __ldcg(src + 2), w3 = __ldcg(src + 3);
      const double4 v0 = make_double4(w0.x, w0.y, w1.x, w1.y), v1 = make_double4(w2.x, w2.y, w3.x, w3.y);
      p[0] += v0.x * xv; p[1] += v0.y * xv; p[2] += v0.z * xv; p[3] += v0.w * xv;
      p[4] += v1.x * xv; p[5] += v1.y * xv; p[6] += v1.z * xv; p[7] += v1.w * xv;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) p[c] += __shfl_xor_sync(0xffffffffu, p[c], d);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) tb[8 * s + c] = p[c];
    }
  }
  // inverses of the diagonal blocks: thread (s, c) solves L_ss y = e_c
  if (t < 8 * KT) {
    const int s = t >> 3, c = t & 7;
    const double *Ld = Ls + ls_off(s);       // rows 8 s .. 8 s + 7 of panel s: L_ss[r][m] at Ld[r * 8 + m]
    double y[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) y[r] = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double l = Ld[r * 8 + r];
      y[r] *= (l > 0.0) ? 1.0 / l : 0.0;     // non-positive pivot: zero component, like LDLT's D^+
#pragma unroll
      for (int q = r + 1; q < 8; ++q) y[q] -= Ld[q * 8 + r] * y[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) Li[s * 64 + r * 8 + c] = y[r];
  }
  __syncthreads();
  // block chain inside warp 0: x_s = -L_ss^-T tb_s, then tb[c] += L[8 s + q][c] x_q for the columns left of it
  if (warp == 0) {
    for (int s = KT - 1; s >= 0; --s) {
      if (lane < 8) {
        double xa = 0.0, xb = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          xa += (i >= lane) ? Li[s * 64 + i * 8 + lane] * tb[8 * s + i] : 0.0;
          xb += (i + 1 >= lane) ? Li[s * 64 + (i + 1) * 8 + lane] * tb[8 * s + i + 1] : 0.0;
        }
        sm.xs[8 * s + lane] = -(xa + xb);
      }
      __syncwarp();
      double xq[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) xq[q] = sm.xs[8 * s + q];
      for (int c = lane; c < 8 * s; c += 32) {
        const double *Lc = Ls + ls_off(c >> 3) + (size_t)(8 * s - 8 * (c >> 3)) * 8 + (c & 7);   // L[8 s + q][c] at Lc[8 q]
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) u += Lc[8 * q] * xq[q];
        tb[c] += u;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int j = t; j < nd.k; j += kNdThreads) g.x[nd.own0 + j] = sm.xs[j];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// drivers
// ------------------------------------------------------------------------------------------------------------
template <int TPW>
__global__ void __launch_bounds__(kNdThreads, 1)
k_nd_forward_level(NdArgs g, int list_begin, int timing, const LmState *st) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char nd_raw[];
  const NdSmem sm = nd_carve(nd_raw, g);
  nd_forward_node<TPW>(g, sm, g.list[list_begin + blockIdx.x], timing);
}

__global__ void __launch_bounds__(kNdThreads, 1)
k_nd_backward_level(NdArgs g, int list_begin, const LmState *st) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char nd_raw[];
  const NdSmem sm = nd_carve(nd_raw, g);
  nd_backward_node(g, sm, g.list[list_begin + blockIdx.x]);
}

__device__ __forceinline__ int nd_ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void nd_st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 spins (bounded: a lost hand-over must not hang the device), everybody follows through the barrier
__device__ __forceinline__ void nd_wait_flag(const NdArgs &g, int idx) {
  if (threadIdx.x == 0) {
    int it = 0;
    while (nd_ld_acquire(g.flags + idx) == 0) {
      if (++it > kNdSpinLimit || nd_ld_acquire(g.flags + 2 * g.n_nodes) != 0) {
        atomicExch(g.flags + 2 * g.n_nodes, 1);       // abort this launch
        atomicExch(g.flags + 2 * g.n_nodes + 1, 1);   // sticky: reported by the host (never cleared by the launch)
        break;
      }
      __nanosleep(40);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void nd_set_flag(const NdArgs &g, int idx) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    nd_st_release(g.flags + idx, 1);
  }
}

// One launch: CTA p eliminates leaf p and climbs while it arrives as child[0]; flags hand the contribution blocks
// (forward) and the solved boundary values (backward) over between CTAs.  The flags are cleared by a memset node
// that precedes the launch.  Requires all CTAs to be co-resident (grid <= SM count, one CTA per SM).
template <int TPW>
__global__ void __launch_bounds__(kNdThreads, 1)
k_nd_persistent(NdArgs g, int timing, const LmState *st) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char nd_raw[];
  const NdSmem sm = nd_carve(nd_raw, g);
  const int lb = g.list_ptr[blockIdx.x], le = g.list_ptr[blockIdx.x + 1];
  unsigned long long t0 = 0;
  if (timing && threadIdx.x == 0) t0 = gtime();
  for (int q = lb; q < le; ++q) {
    const int id = g.list[q];
    const int other = g.nodes[id].child[1];
    if (other >= 0) nd_wait_flag(g, other);
    nd_forward_node<TPW>(g, sm, id, timing);
    nd_set_flag(g, id);
  }
  if (timing && threadIdx.x == 0 && blockIdx.x == 0) g_nd_dbg[0] = gtime() - t0;
  for (int q = le - 1; q >= lb; --q) {
    const int id = g.list[q];
    const int parent = g.nodes[id].parent;
    if (q == le - 1 && parent >= 0) nd_wait_flag(g, g.n_nodes + parent);   // further down the list the parent is mine
    nd_backward_node(g, sm, id);
    nd_set_flag(g, g.n_nodes + id);
  }
  if (timing && threadIdx.x == 0 && blockIdx.x == 0) g_nd_dbg[1] = gtime() - t0;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct NdDevice {
  NdArgs level_args{}, cta_args{};
  size_t smem = 0;
  int tpw = 0;
};

inline int nd_tpw_for(int BT) {
  const int need = (BT * (BT + 1) / 2 + kNdCons - 1) / kNdCons;
  const int opts[] = {4, 8, 12, 17, 24, 29};
  for (int o : opts)
    if (need <= o) return o;
  return -1;
}

template <int TPW>
inline bool nd_enqueue_t(const NdPlan &pl, const NdDevice &dv, const double *Saug, double *x, int mode, int timing,
                         const LmState *st, cudaStream_t stream, long long *launches) {
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    cudaFuncSetAttribute(k_nd_forward_level<TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNdSmemLimit);
    cudaFuncSetAttribute(k_nd_persistent<TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNdSmemLimit);
    cudaFuncSetAttribute(k_nd_backward_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNdSmemLimit);
    attr_done[dev] = true;
  }
  if (mode == 6) {
    cudaMemsetAsync(dv.cta_args.flags, 0, (size_t)(2 * dv.cta_args.n_nodes + 1) * sizeof(int), stream);
    NdArgs a = dv.cta_args;
    a.S = Saug; a.x = x;
    void *args[] = {(void *)&a, (void *)&timing, (void *)&st};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.n_ctas);
    cfg.blockDim = dim3(kNdThreads);
    cfg.dynamicSmemBytes = dv.smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)k_nd_persistent<TPW>, args);
    if (e != cudaSuccess) {
      cudaGetLastError();
      k_nd_persistent<TPW><<<pl.n_ctas, kNdThreads, dv.smem, stream>>>(a, timing, st);
      e = cudaGetLastError();
    }
    if (launches) *launches += 1;
    return e == cudaSuccess;
  }
  NdArgs la = dv.level_args;
  la.S = Saug; la.x = x;
  for (int l = 0; l < pl.n_levels; ++l) {
    const int cnt = pl.level_ptr[l + 1] - pl.level_ptr[l];
    k_nd_forward_level<TPW><<<cnt, kNdThreads, dv.smem, stream>>>(la, pl.level_ptr[l], timing, st);
  }
  for (int l = pl.n_levels - 1; l >= 0; --l) {
    const int cnt = pl.level_ptr[l + 1] - pl.level_ptr[l];
    k_nd_backward_level<<<cnt, kNdThreads, dv.smem, stream>>>(la, pl.level_ptr[l], st);
  }
  if (launches) *launches += 2 * pl.n_levels;
  return cudaGetLastError() == cudaSuccess;
}

// mode 5: one launch per level; mode 6: one persistent launch
inline bool nd_enqueue(const NdPlan &pl, const NdDevice &dv, const double *Saug, double *x, int mode, int timing,
                       const LmState *st, cudaStream_t stream, long long *launches) {
  switch (dv.tpw) {
    case 4: return nd_enqueue_t<4>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    case 8: return nd_enqueue_t<8>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    case 12: return nd_enqueue_t<12>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    case 17: return nd_enqueue_t<17>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    case 24: return nd_enqueue_t<24>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    case 29: return nd_enqueue_t<29>(pl, dv, Saug, x, mode, timing, st, stream, launches);
    default: return false;
  }
}

}  // namespace ba
