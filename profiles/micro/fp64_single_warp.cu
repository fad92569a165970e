// How fast can ONE warp issue FP64 FMAs on B200, and do warps on the same SM sub-partition (warps w, w+4, w+8 of
// a CTA) share that rate?  Answers what bounds the single-warp pivot chain of the banded Cholesky.
//   nvcc -arch=sm_100a -O3 -o fp64_single_warp fp64_single_warp.cu && ./fp64_single_warp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, unsigned warp_mask) {
  const int warp = threadIdx.x >> 5;
  if (!((warp_mask >> warp) & 1u)) return;
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

template <int ILP>
void run(const char *name, unsigned mask, double *out, long long *cyc) {
  const int iters = 4096;
  k<ILP><<<1, 384>>>(out, cyc, iters, mask);
  k<ILP><<<1, 384>>>(out, cyc, iters, mask);
  cudaDeviceSynchronize();
  long long h[12];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  int first = 0;
  while (!((mask >> first) & 1u)) ++first;
  printf("%-28s ILP %d: %.2f cycles per DFMA per warp\n", name, ILP, (double)h[first] / (iters * (double)ILP));
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 384 * sizeof(double));
  cudaMalloc(&cyc, 12 * sizeof(long long));
  run<1>("one warp", 1u << 3, out, cyc);
  run<2>("one warp", 1u << 3, out, cyc);
  run<4>("one warp", 1u << 3, out, cyc);
  run<8>("one warp", 1u << 3, out, cyc);
  run<8>("warps 3,7 (same SMSP)", (1u << 3) | (1u << 7), out, cyc);
  run<8>("warps 3,7,11 (same SMSP)", (1u << 3) | (1u << 7) | (1u << 11), out, cyc);
  run<8>("warps 0,1,2,3 (4 SMSPs)", 0xfu, out, cyc);
  run<1>("warps 3,7,11 (same SMSP)", (1u << 3) | (1u << 7) | (1u << 11), out, cyc);
  run<8>("all 12 warps", 0xfffu, out, cyc);
  return 0;
}
