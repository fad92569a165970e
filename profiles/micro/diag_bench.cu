// Microbenchmark of the 64x64 diagonal-block factorisation variants (cycles per block, one CTA).
#include <cstdio>
#include <vector>
#include <cmath>
#define BA_DIAG_TIMING 1
__device__ long long g_t[8];
#include "../../bundle_adjustment_solver_b200/csrc/ba_cholesky_cluster.cuh"
using namespace ba;

template <int VARIANT>
__global__ void __launch_bounds__(256) k_bench(double *A, int ld, int reps, long long *cyc, const double *A0) {
  extern __shared__ double sm[];
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = threadIdx.x; e < 64 * 65; e += 256) A[e] = A0[e];
    __syncthreads();
    long long t0 = clock64();
    if (VARIANT == 0) chol_panel_tall<0>(A, ld, 65, 0, 64, nullptr, sm);
    else { __shared__ int rw[3]; rw[0] = 1; rw[1] = 2; rw[2] = 3; chol_panel_tall<3>(A, ld, 64, 0, 64, rw, sm); }
    __syncthreads();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) *cyc = total / reps;
}

int main() {
  const int n = 64, ld = 65;
  std::vector<double> M(n * n), S(ld * ld, 0.0);
  srand(1);
  for (auto &v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c <= r; ++c) {
      double s = (r == c) ? 5.0 : 0.0;
      for (int k = 0; k < n; ++k) s += M[r * n + k] * M[c * n + k];
      S[c * ld + r] = s;
    }
  double *dA, *dA0; long long *dc;
  cudaMalloc(&dA, sizeof(double) * ld * ld); cudaMalloc(&dA0, sizeof(double) * ld * ld); cudaMalloc(&dc, 8);
  cudaMemcpy(dA0, S.data(), sizeof(double) * ld * ld, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem);
  k_bench<0><<<1, 256, kClusterSmem>>>(dA, ld, 20, dc, dA0);
  long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  printf("chol_diag_2d: %lld cycles per 64x64 block (%.1f per column) err=%s\n", h, h / 64.0, cudaGetErrorString(cudaGetLastError()));
  // verify
  std::vector<double> L(ld * ld);
  cudaMemcpy(L.data(), dA, sizeof(double) * ld * ld, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c <= r; ++c) {
      double s = 0;
      for (int k = 0; k <= c; ++k) s += L[k * ld + r] * L[k * ld + c];
      maxerr = fmax(maxerr, fabs(s - S[c * ld + r]));
    }
  printf("max |LL^T - S| = %g\n", maxerr);
  long long ht[8]; cudaMemcpyFromSymbol(ht, g_t, sizeof(ht));
  printf("sections: load %lld loop %lld sqrt %lld scale+store %lld\n", ht[0], ht[1], ht[2], ht[3]);
  cudaFuncSetAttribute(k_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem);
  k_bench<1><<<1, 256, kClusterSmem>>>(dA, ld, 20, dc, dA0);
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpyFromSymbol(ht, g_t, sizeof(ht));
  printf("tall<3> (timing only, rows masked): %lld cycles; loop %lld (%.1f per column) %s\n", h, ht[1], ht[1] / 64.0, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
