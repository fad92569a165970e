// Variants of the 64x64 column-sweep to find what bounds the per-column latency.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
constexpr int kCB = 64, kCLD = 65;

__device__ __forceinline__ double rcp_f32newton(double d) {
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__double2float_rn(d)));
  double x = (double)r; double e = fma(-d, x, 1.0); x = fma(x, e, x); e = fma(-d, x, 1.0); x = fma(x, e, x); return x;
}
__device__ __forceinline__ double rcp_f64approx(double d) {
  double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0); x = fma(x, e, x); e = fma(-d, x, 1.0); x = fma(x, e, x); return x;
}

// RCP: 0 = 1.0/d, 1 = f32 newton, 2 = f64 approx newton, 3 = no reciprocal at all (di = 1e-3 constant; wrong math, timing only)
// NOSTORE: skip Lraw store ; NOBAR: skip barrier (wrong, timing only)
template <int RCP, bool NOSTORE, bool NOBAR, bool NOUPD>
__device__ __forceinline__ long long sweep(double *A, int ld, double *sm) {
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  double (*Lraw)[kCLD] = reinterpret_cast<double (*)[kCLD]>(sm);
  double *cb = sm + 2 * kCB * kCLD;
  double *dib = cb + 2 * kCB;
  double v[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = ty + 16 * a, c = tx + 16 * b;
      v[a][b] = (c <= r) ? A[(size_t)c * ld + r] : 0.0;
    }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll
  for (int j = 0; j < kCB; ++j) {
    const int par = (j & 1) * kCB, g = j >> 4, jm = j & 15;
    if (tx == jm) {
#pragma unroll
      for (int a = g; a < 4; ++a) {
        const int r = ty + 16 * a;
        if (r >= j) { cb[par + r] = v[a][g]; if (!NOSTORE) Lraw[r][j] = v[a][g]; }
      }
      if (ty == jm) {
        const double d = v[g][g];
        double di;
        if (RCP == 0) di = 1.0 / d; else if (RCP == 1) di = rcp_f32newton(d); else if (RCP == 2) di = rcp_f64approx(d); else di = 1e-3;
        dib[j & 1] = di;
      }
    }
    if (!NOBAR) __syncthreads();
    const double di = dib[j & 1];
    double lr[4], lc[4];
#pragma unroll
    for (int a = g; a < 4; ++a) lr[a] = cb[par + ty + 16 * a] * di;
#pragma unroll
    for (int b = g; b < 4; ++b) lc[b] = cb[par + tx + 16 * b];
    if (!NOUPD) {
#pragma unroll
      for (int a = g; a < 4; ++a)
#pragma unroll
        for (int b = g; b <= a; ++b) v[a][b] -= lr[a] * lc[b];
    } else {
      v[3][3] -= lr[3] * lc[3];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) s += v[a][b];
  A[t] = s;  // keep alive
  return t1 - t0;
}

template <int RCP, bool NOSTORE, bool NOBAR, bool NOUPD>
__global__ void __launch_bounds__(256) k_bench(double *A, int ld, int reps, long long *cyc, const double *A0) {
  extern __shared__ double sm[];
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = threadIdx.x; e < 64 * 65; e += 256) A[e] = A0[e];
    __syncthreads();
    total += sweep<RCP, NOSTORE, NOBAR, NOUPD>(A, ld, sm);
    __syncthreads();
  }
  if (threadIdx.x == 0) *cyc = total / reps;
}

template <int RCP, bool NOSTORE, bool NOBAR, bool NOUPD>
void run(const char *name, double *dA, double *dA0, long long *dc) {
  const size_t smem = (2 * kCB * kCLD + 8 * kCB) * sizeof(double);
  cudaFuncSetAttribute(k_bench<RCP, NOSTORE, NOBAR, NOUPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_bench<RCP, NOSTORE, NOBAR, NOUPD><<<1, 256, smem>>>(dA, 65, 20, dc, dA0);
  long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s loop %6lld cycles = %.1f per column (%s)\n", name, h, h / 64.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = 64, ld = 65;
  std::vector<double> M(n * n), S(ld * ld, 0.0);
  srand(1);
  for (auto &v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < n; ++r) for (int c = 0; c <= r; ++c) { double s = (r == c) ? 5.0 : 0.0; for (int k = 0; k < n; ++k) s += M[r * n + k] * M[c * n + k]; S[c * ld + r] = s; }
  double *dA, *dA0; long long *dc;
  cudaMalloc(&dA, sizeof(double) * ld * ld); cudaMalloc(&dA0, sizeof(double) * ld * ld); cudaMalloc(&dc, 8);
  cudaMemcpy(dA0, S.data(), sizeof(double) * ld * ld, cudaMemcpyHostToDevice);
  run<0, false, false, false>("rcp=1.0/d", dA, dA0, dc);
  run<1, false, false, false>("rcp=f32+2newton", dA, dA0, dc);
  run<2, false, false, false>("rcp=f64approx+2newton", dA, dA0, dc);
  run<3, false, false, false>("rcp=none (timing only)", dA, dA0, dc);
  run<3, true, false, false>("rcp=none, no Lraw store", dA, dA0, dc);
  run<3, true, true, false>("rcp=none, no store, no barrier", dA, dA0, dc);
  run<3, true, false, true>("rcp=none, no store, 1 update only", dA, dA0, dc);
  run<2, true, false, false>("rcp=f64approx, no Lraw store", dA, dA0, dc);
  return 0;
}
