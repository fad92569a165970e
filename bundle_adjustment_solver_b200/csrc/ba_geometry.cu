// ba_geometry.cu -- batched device versions of the reference's SE(3) / SO(3) / quaternion / Euler helpers
// (utility/geometry_library.cpp:93-736), one thread per element.  The arithmetic is the shared host/device source
// include/ba_b200/utility/geometry_math.h; this file only adds the kernel and the C-ABI entry points.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/ba_b200.h"
#include "../../include/ba_b200/utility/geometry_math.h"

namespace {

struct GeomShape { int in, in2, out; };
__host__ __device__ inline GeomShape geom_shape(int op) {
  switch (op) {
    case BA_GEOM_SE3_EXP: return {6, 0, 12};
    case BA_GEOM_SE3_LOG: return {12, 0, 6};
    case BA_GEOM_SO3_EXP: return {3, 0, 9};
    case BA_GEOM_SO3_LOG: return {9, 0, 3};
    case BA_GEOM_Q2R: return {4, 0, 9};
    case BA_GEOM_R2Q: return {9, 0, 4};
    case BA_GEOM_ROTVEC2Q: return {3, 0, 4};
    case BA_GEOM_R2EULER: return {9, 0, 3};
    case BA_GEOM_A2R: return {3, 0, 9};
    case BA_GEOM_INVERSE_SE3: return {12, 0, 12};
    case BA_GEOM_ADD_FRONT_SE3: return {6, 6, 6};
    case BA_GEOM_Q_MULT: return {4, 4, 4};
    default: return {0, 0, 0};
  }
}

template <typename T>
__global__ void k_geometry(int op, long long n, const T *__restrict__ in, const T *__restrict__ in2, T *__restrict__ out) {
  const GeomShape sh = geom_shape(op);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T a[12], b[6], o[12];
    for (int k = 0; k < sh.in; ++k) a[k] = in[i * sh.in + k];
    for (int k = 0; k < sh.in2; ++k) b[k] = in2[i * sh.in2 + k];
    switch (op) {
      case BA_GEOM_SE3_EXP: ba_geom::se3_exp(a, o, o + 9); break;
      case BA_GEOM_SE3_LOG: ba_geom::se3_log(a, a + 9, o); break;
      case BA_GEOM_SO3_EXP: ba_geom::so3_exp(a, o); break;
      case BA_GEOM_SO3_LOG: ba_geom::so3_log(a, o); break;
      case BA_GEOM_Q2R: ba_geom::q2r(a, o); break;
      case BA_GEOM_R2Q: ba_geom::r2q(a, o); break;
      case BA_GEOM_ROTVEC2Q: ba_geom::rotvec2q(a, o); break;
      case BA_GEOM_R2EULER: ba_geom::r2euler(a, o); break;
      case BA_GEOM_A2R: ba_geom::a2r(a, o); break;
      case BA_GEOM_INVERSE_SE3: ba_geom::inverse_se3(a, a + 9, o, o + 9); break;
      case BA_GEOM_ADD_FRONT_SE3: ba_geom::add_front_se3(a, b, o); break;
      case BA_GEOM_Q_MULT: ba_geom::q_mult(a, b, o); break;
      default: break;
    }
    for (int k = 0; k < sh.out; ++k) out[i * sh.out + k] = o[k];
  }
}

template <typename T>
int geometry_batched(int device, int op, long long n, const T *in, const T *in2, T *out) {
  const GeomShape sh = geom_shape(op);
  if (sh.in == 0 || n < 0 || (n > 0 && (!in || !out)) || (sh.in2 > 0 && n > 0 && !in2)) return BA_ERR_INVALID;
  if (n == 0) return BA_OK;
  if (cudaSetDevice(device) != cudaSuccess) {
    fprintf(stderr, "[ba_b200] ba_geometry_batched: no CUDA device %d (there is no CPU fallback)\n", device);
    cudaGetLastError();
    return BA_ERR_CUDA;
  }
  T *d_in = nullptr, *d_in2 = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_in, sizeof(T) * n * sh.in);
  if (e == cudaSuccess && sh.in2) e = cudaMalloc(&d_in2, sizeof(T) * n * sh.in2);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(T) * n * sh.out);
  if (e == cudaSuccess) e = cudaMemcpy(d_in, in, sizeof(T) * n * sh.in, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && sh.in2) e = cudaMemcpy(d_in2, in2, sizeof(T) * n * sh.in2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_geometry<T><<<grid, 256>>>(op, n, d_in, d_in2, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(T) * n * sh.out, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_in2); cudaFree(d_out);
  if (e != cudaSuccess) {
    fprintf(stderr, "[ba_b200] ba_geometry_batched: %s\n", cudaGetErrorString(e));
    cudaGetLastError();
    return BA_ERR_CUDA;
  }
  return BA_OK;
}

}  // namespace

extern "C" {
int ba_geometry_batched(int device, int op, long long n, const double *in, const double *in2, double *out) {
  return geometry_batched<double>(device, op, n, in, in2, out);
}
int ba_geometry_batched_f(int device, int op, long long n, const float *in, const float *in2, float *out) {
  return geometry_batched<float>(device, op, n, in, in2, out);
}
}
