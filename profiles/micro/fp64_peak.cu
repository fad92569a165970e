// FP64 FMA throughput and dependent-issue latency microbenchmark (yardstick for K4/K5 design).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_fma(double *out, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = x * a + b;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sync(long long *cyc, int iters) {
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  if (threadIdx.x == 0) *cyc = clock64() - t0;
}
template <int ILP>
float run(int blocks, int threads, int iters) {
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k_fma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    float ms = run<8>(p.multiProcessorCount * 2, threads, iters);
    double fma = (double)p.multiProcessorCount * 2 * threads * iters * 8;
    printf("threads=%4d ILP=8: %.3f ms -> %.2f TFLOP/s FP64 (%.1f FMA/clk/SM at 1.9 GHz)\n", threads, ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
  }
  double *out; long long *cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  long long h;
  k_lat<<<1, 32>>>(out, cyc, 10000, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.1f cycles\n", h / 10000.0);
  for (int threads : {64, 256, 1024}) {
    k_sync<<<1, threads>>>(cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads with %d threads: %.1f cycles\n", threads, h / 10000.0);
  }
  return 0;
}
