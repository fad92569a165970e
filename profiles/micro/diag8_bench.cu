// Microbenchmark: one warp factors an 8 x 8 SPD block redundantly in every lane and inverts the factor (the diagonal
// warp of csrc/ba_cholesky_nd.cuh).  Cycles per block for variants of the chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o diag8_bench diag8_bench.cu && ./diag8_bench
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
// one Newton step less: the hardware seed has ~2^-22 relative error -> 2^-44 after one step, 2^-88 after two
__device__ __forceinline__ double rcp1(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// VAR 0: factor (rcp chain) + rsqrt + L + W + publish to shared memory (the kernel's code)
// VAR 1: VAR 0 without the publish (a checksum keeps the values alive)
// VAR 2: factor only
// VAR 3: factor with the rsqrt ON the chain (columns scaled by rsqrt, no rcp), L + W, no publish
// VAR 4: as VAR 0 but W by rows (row i of W from the rows above), publish
// VAR 5: 2 x 2 block pivots (one reciprocal of the determinant per two columns), L via LDL^T, W, publish
template <int VAR>
__global__ void k_bench(const double *A0, double *out, int reps, long long *cyc) {
  __shared__ double tile[64], Lt[64], Wt[64];
  const int lane = threadIdx.x;
  const int fr = lane >> 2, fc = 2 * (lane & 3);
  long long total = 0;
  double chk = 0.0;
  for (int r = 0; r < reps; ++r) {
    for (int e = lane; e < 64; e += 32) tile[e] = A0[e] + 1e-9 * r;
    __syncwarp();
    const long long t0 = clock64();
    double D[8][8], rc[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) D[i][j] = tile[i * 8 + j];
    double Lm[8][8], W[8][8];
    if (VAR == 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        rs[k] = fast_rsqrt(D[k][k]);
#pragma unroll
        for (int i = k; i < 8; ++i) D[i][k] *= rs[k];
#pragma unroll
        for (int i = k + 1; i < 8; ++i)
#pragma unroll
          for (int j = k + 1; j <= i; ++j) D[i][j] -= D[i][k] * D[j][k];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) Lm[i][j] = (j <= i) ? D[i][j] : 0.0;
    } else if (VAR == 5) {
      // LDL^T with 2 x 2 diagonal blocks: E = [[p, q], [q, r]], E^-1 = [[r, -q], [-q, p]] / (p r - q^2)
      double dinv[8];   // reciprocal pivots of the scalar LDL^T recovered per block for the final scaling
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const double p = D[k][k], q = D[k + 1][k], rr = D[k + 1][k + 1];
        const double det = fma(p, rr, -q * q);
        const double idet = fast_rcp(det);
        const double e00 = rr * idet, e01 = -q * idet, e11 = p * idet;
        // scalar pivots: d0 = p, d1 = r - q^2 / p = det / p
        dinv[k] = 0.0; dinv[k + 1] = 0.0;   // filled below (off the chain)
#pragma unroll
        for (int i = k + 2; i < 8; ++i) {
          const double u0 = D[i][k] * e00 + D[i][k + 1] * e01;    // row i of  A21 E^-1
          const double u1 = D[i][k] * e01 + D[i][k + 1] * e11;
#pragma unroll
          for (int j = k + 2; j <= i; ++j) D[i][j] -= u0 * D[j][k] + u1 * D[j][k + 1];
        }
      }
      // scalar Cholesky of the 2 x 2 diagonal blocks and of the columns below (off the chain: needs only the D
      // entries each block step left behind): column k: L[i][k] = D[i][k] / sqrt(p); column k+1 from the 2 x 2 update
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const double p = D[k][k], q = D[k + 1][k];
        rs[k] = fast_rsqrt(p);
        const double l10 = q * rs[k];
        const double d1 = fma(-l10, l10, D[k + 1][k + 1]);
        rs[k + 1] = fast_rsqrt(d1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          Lm[i][k] = (i >= k) ? D[i][k] * rs[k] : 0.0;
          Lm[i][k + 1] = (i >= k + 1) ? ((i == k + 1) ? d1 * rs[k + 1] : fma(-Lm[i][k], l10, D[i][k + 1]) * rs[k + 1]) : 0.0;
        }
      }
      (void)dinv;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double d = D[k][k];
        const bool pos_def = d > 0.0;
        rc[k] = pos_def ? fast_rcp(d) : 0.0;
        if (VAR != 2) rs[k] = pos_def ? fast_rsqrt(d) : 0.0;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) {
          const double ti = D[i][k] * rc[k];
#pragma unroll
          for (int j = k + 1; j <= i; ++j) D[i][j] -= ti * D[j][k];
        }
      }
      if (VAR != 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) Lm[i][j] = (j <= i) ? D[i][j] * rs[j] : 0.0;
      }
    }
    if (VAR == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) chk += D[i][i] + D[7][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) W[i][j] = 0.0;
      if (VAR == 4) {
        // row i of W: W[i][j] = -rs_i sum_{m=j}^{i-1} L[i][m] W[m][j]  (all j of a row are independent)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          W[i][i] = rs[i];
#pragma unroll
          for (int j = 0; j < i; ++j) {
            double sacc = 0.0;
#pragma unroll
            for (int m = j; m < i; ++m) sacc += Lm[i][m] * W[m][j];
            W[i][j] = -sacc * rs[i];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          W[j][j] = rs[j];
#pragma unroll
          for (int i = j + 1; i < 8; ++i) {
            double sacc = 0.0;
#pragma unroll
            for (int m = j; m < i; ++m) sacc += Lm[i][m] * W[m][j];
            W[i][j] = -sacc * rs[i];
          }
        }
      }
      if (VAR == 1 || VAR == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) chk += Lm[i][j] + W[i][j];
      } else {
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          if (fr == rr) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              if (fc == c) {
                *reinterpret_cast<double2 *>(Lt + rr * 8 + c) = make_double2(Lm[rr][c], Lm[rr][c + 1]);
                *reinterpret_cast<double2 *>(Wt + rr * 8 + c) = make_double2(W[rr][c], W[rr][c + 1]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    total += clock64() - t0;
  }
  if (lane == 0) *cyc = total / reps;
  if (VAR == 1 || VAR == 2 || VAR == 3) { if (lane == 0) out[0] = chk; }
  else for (int e = lane; e < 64; e += 32) { out[e] = Lt[e]; out[64 + e] = Wt[e]; }
}

template <int VAR>
void run(const double *dA, double *dout, long long *dc, const std::vector<double> &S, const char *name) {
  k_bench<VAR><<<1, 32>>>(dA, dout, 50, dc);
  long long h = 0;
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  std::vector<double> o(128);
  cudaMemcpy(o.data(), dout, 128 * 8, cudaMemcpyDeviceToHost);
  double err = -1.0;
  if (VAR == 0 || VAR == 4 || VAR == 5) {
    err = 0.0;
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c <= r; ++c) {
        double s = 0.0, w = 0.0;
        for (int k = 0; k <= c; ++k) s += o[r * 8 + k] * o[c * 8 + k];
        for (int k = c; k <= r; ++k) w += o[64 + r * 8 + k] * o[k * 8 + c];   // (W L)[r][c]
        err = fmax(err, fabs(s - S[r * 8 + c]));
        err = fmax(err, fabs(w - (r == c ? 1.0 : 0.0)));
      }
  }
  printf("%-70s %6lld cycles per block   max err %.2e  (%s)\n", name, h, err, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  std::vector<double> M(64), S(64, 0.0);
  srand(3);
  for (auto &v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) {
      double s = (r == c) ? 2.0 : 0.0;
      for (int k = 0; k < 8; ++k) s += M[r * 8 + k] * M[c * 8 + k];
      S[r * 8 + c] = s;
    }
  double *dA, *dout;
  long long *dc;
  cudaMalloc(&dA, 64 * 8); cudaMalloc(&dout, 128 * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, S.data(), 64 * 8, cudaMemcpyHostToDevice);
  run<0>(dA, dout, dc, S, "0: rcp chain + rsqrt + L + W (columns) + publish");
  run<1>(dA, dout, dc, S, "1: as 0, no publish");
  run<2>(dA, dout, dc, S, "2: rcp chain only");
  run<3>(dA, dout, dc, S, "3: rsqrt on the chain + W, no publish");
  run<4>(dA, dout, dc, S, "4: as 0, W by rows");
  run<5>(dA, dout, dc, S, "5: 2x2 block pivots + L + W + publish");
  return 0;
}
