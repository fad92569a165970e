// DMMA (mma.sync.m8n8k4.f64) throughput / latency microbenchmark on B200: the yardstick for using FP64 tensor
// cores in the Schur accumulation (K4) and the Cholesky trailing update (K5).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b) {
  double c0 = threadIdx.x, c1 = 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
float run(int blocks, int threads, int iters) {
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k_dmma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    float ms = run<8>(p.multiProcessorCount, threads, iters);
    double mmas = (double)p.multiProcessorCount * (threads / 32) * iters * 8;
    printf("threads=%4d ILP=8: %.3f ms -> %.2f TFLOP/s FP64 DMMA (%.2f cycles per mma per SMSP at 1.965 GHz)\n", threads, ms,
           mmas * 512 / ms / 1e9, ms * 1e-3 * 1.965e9 / (mmas / p.multiProcessorCount / 4));
  }
  double *out; long long *cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  long long h;
  k_lat<<<1, 32>>>(out, cyc, 10000, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DMMA latency: %.1f cycles\n", h / 10000.0);
  return 0;
}
