// ba_cholesky.cuh -- K5: dense FP64 Cholesky solve of the reduced camera system
// (replaces `Am_BCinvBt_mat.ldlt().solve(am_BCinv_b_mat)`, full...cpp:890-908).
//
// Storage: one buffer Saug of (n+1) x (n+1) doubles, leading dimension ld = n+1, viewed
// column-major with the LOWER triangle holding S (equivalently row-major upper, which is how the
// Schur kernel addresses it) and row n (elements c*ld + n) holding rhs^T.  Factorising the leading
// n x n block with the panel TRSM applied to every row below the diagonal block -- including row
// n -- turns that row into z^T = (L^-1 rhs)^T, so the forward substitution is free.  The backward
// substitution L^T x = z runs as a right-looking blocked sweep.
//
// v1: blocked right-looking, NB = 64, three kernels per panel (diag POTRF in one CTA, row-parallel
// TRSM, shared-memory tiled SYRK/GEMM trailing update on CUDA-core FP64).
#pragma once
#include <cuda_runtime.h>

#include "ba_device.cuh"

namespace ba {

constexpr int kNB = 64;

// ---- diagonal block: unblocked Cholesky of an nb x nb block in shared memory ---------------
__global__ void __launch_bounds__(kNB) k_potrf_diag(double *__restrict__ A, int ld, int k0, int nb,
                                                    const LmState *st) {
  if (st->done) return;
  __shared__ double L[kNB][kNB + 1];
  const int t = threadIdx.x;
  // load lower triangle; pad with identity
  for (int c = 0; c < kNB; ++c) {
    double v = (t == c) ? 1.0 : 0.0;
    if (t < nb && c < nb && t >= c) v = A[(size_t)(k0 + c) * ld + k0 + t];
    L[t][c] = v;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    // left-looking column j: L[t][j] = (A[t][j] - sum_{m<j} L[t][m] L[j][m]) / L[j][j]
    double acc = L[t][j];
    if (t >= j) {
      for (int m = 0; m < j; ++m) acc -= L[t][m] * L[j][m];
    }
    __syncthreads();
    // non-positive pivot (e.g. a pose without observations): emulate LDLT's D^+ = 0 by an
    // infinite diagonal, which zeroes the column and the solution component
    if (t == j) L[j][j] = (acc > 0.0) ? sqrt(acc) : __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();
    if (t > j) L[t][j] = acc / L[j][j];
    __syncthreads();
  }
  if (t < nb)
    for (int c = 0; c <= t; ++c) A[(size_t)(k0 + c) * ld + k0 + t] = L[t][c];
}

// ---- panel: rows r in (k0+nb, n] : A[r, k0:k0+nb] <- A[r, k0:k0+nb] L_kk^-T --------------------
__global__ void __launch_bounds__(128) k_trsm_panel(double *__restrict__ A, int ld, int n_rows, int k0,
                                                    int nb, const LmState *st) {
  if (st->done) return;
  __shared__ double L[kNB][kNB + 1];
  __shared__ double invd[kNB];
  for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
    const int r = e % kNB, c = e / kNB;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c < nb && r >= c) v = A[(size_t)(k0 + c) * ld + k0 + r];
    L[r][c] = v;
    if (r == c) invd[r] = 1.0 / v;
  }
  __syncthreads();
  const int r = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  double x[kNB];
#pragma unroll
  for (int c = 0; c < kNB; ++c) x[c] = (c < nb) ? A[(size_t)(k0 + c) * ld + r] : 0.0;
#pragma unroll
  for (int c = 0; c < kNB; ++c) {
    double acc = x[c];
#pragma unroll
    for (int m = 0; m < c; ++m) acc -= x[m] * L[c][m];
    x[c] = acc * invd[c];
  }
#pragma unroll
  for (int c = 0; c < kNB; ++c)
    if (c < nb) A[(size_t)(k0 + c) * ld + r] = x[c];
}

// ---- trailing update: A[r, c] -= sum_m P[r, m] P[c, m], r >= c, both in (k0+nb, n_rows) ------
// 64 x 64 output tile per CTA, 256 threads, 4 x 4 per thread, K = nb <= 64 staged in smem.
__global__ void __launch_bounds__(256) k_syrk_update(double *__restrict__ A, int ld, int n_rows, int k0,
                                                     int nb, const LmState *st) {
  if (st->done) return;
  const int base = k0 + nb;
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc > tr) return;
  constexpr int KH = 32;                // K staged in halves to stay under the 48 KB static limit
  __shared__ double Pr[KH][64 + 1];     // [m][row]
  __shared__ double Pc[KH][64 + 1];     // [m][col]
  const int r0 = base + tr * 64, c0 = base + tc * 64;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int mh = 0; mh < nb; mh += KH) {
    __syncthreads();
    for (int e = threadIdx.x; e < KH * 64; e += 256) {
      const int i = e % 64, m = mh + e / 64;
      const int rr = r0 + i, cc = c0 + i;
      Pr[e / 64][i] = (m < nb && rr < n_rows) ? A[(size_t)(k0 + m) * ld + rr] : 0.0;
      Pc[e / 64][i] = (m < nb && cc < n_rows) ? A[(size_t)(k0 + m) * ld + cc] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < KH; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Pr[m][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Pc[m][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + ty + 16 * j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = r0 + tx + 16 * i;
      if (rr < n_rows && cc < n_rows - 1 && rr >= cc) A[(size_t)cc * ld + rr] -= acc[i][j];
    }
  }
}

// ---- backward substitution L^T x = z, z = row n of the factor; single CTA (v1) -------------------
// x_out[0..n).  Right-looking: after block k is solved, z[0:k0] -= L[k0:k0+nb, 0:k0]^T x_k.
__global__ void __launch_bounds__(1024) k_backward_solve(const double *__restrict__ A, int ld, int n,
                                                         double *__restrict__ x_out, double *__restrict__ zbuf,
                                                         const LmState *st) {
  if (st->done) return;
  __shared__ double xk[kNB];
  __shared__ double Ld[kNB][kNB + 1];
  const int t = threadIdx.x;
  for (int i = t; i < n; i += blockDim.x) zbuf[i] = A[(size_t)i * ld + n];
  __syncthreads();
  const int nblk = (n + kNB - 1) / kNB;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int k0 = kb * kNB, nb = min(kNB, n - k0);
    for (int e = t; e < kNB * kNB; e += blockDim.x) {
      const int r = e % kNB, c = e / kNB;
      Ld[r][c] = (r < nb && c < nb && r >= c) ? A[(size_t)(k0 + c) * ld + k0 + r] : ((r == c) ? 1.0 : 0.0);
    }
    if (t < kNB) xk[t] = (t < nb) ? zbuf[k0 + t] : 0.0;
    __syncthreads();
    if (t < 32) {
      // solve L_kk^T x = z backwards; lane owns entries t and t+32
      for (int j = nb - 1; j >= 0; --j) {
        const double xj = xk[j] / Ld[j][j];
        __syncwarp();
        if (t == 0) xk[j] = xj;
        for (int i = t; i < j; i += 32) xk[i] -= Ld[j][i] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
    if (t < nb) x_out[k0 + t] = xk[t];
    // z[c] -= sum_r L[k0+r, c] * x_k[r], c < k0 ; warp per column, lanes over rows
    const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    for (int c = warp; c < k0; c += nwarps) {
      double acc = 0.0;
      for (int r = lane; r < nb; r += 32) acc += A[(size_t)c * ld + k0 + r] * xk[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
      if (lane == 0) zbuf[c] -= acc;
    }
    __syncthreads();
  }
}

struct CholeskyPlan {
  int n = 0;
};

// Enqueue factor + solve on `stream`.  Saug: (n+1)^2 doubles, ld = n+1.  x: n doubles.  zbuf: n.
inline void cholesky_solve_enqueue(double *Saug, int n, double *x, double *zbuf, const LmState *st,
                                   cudaStream_t stream, long long *launches) {
  const int ld = n + 1, n_rows = n + 1;
  for (int k0 = 0; k0 < n; k0 += kNB) {
    const int nb = (n - k0 < kNB) ? (n - k0) : kNB;
    k_potrf_diag<<<1, kNB, 0, stream>>>(Saug, ld, k0, nb, st);
    const int rows_below = n_rows - (k0 + nb);
    if (rows_below > 0) {
      k_trsm_panel<<<(rows_below + 127) / 128, 128, 0, stream>>>(Saug, ld, n_rows, k0, nb, st);
      const int tiles = (rows_below + 63) / 64;
      dim3 g(tiles, tiles);
      k_syrk_update<<<g, 256, 0, stream>>>(Saug, ld, n_rows, k0, nb, st);
      if (launches) *launches += 2;
    }
    if (launches) *launches += 1;
  }
  k_backward_solve<<<1, 1024, 0, stream>>>(Saug, ld, n, x, zbuf, st);
  if (launches) *launches += 1;
}

}  // namespace ba
