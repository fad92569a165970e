// ba_poseonly.cu -- K8: batched pose-only bundle adjustment (FP32), one thread group per frame,
// persistent over the damped Gauss-Newton iterations.
// Reference: core/pose_only_bundle_adjustment_solver.cpp
//   Solve_Monocular_6Dof :8-170, Solve_Stereo_6Dof :172-399,
//   Solve_Monocular_Planar3Dof :401-615, Solve_Stereo_Planar3Dof :617-900,
//   Jacobians :1350-1384 / :1454-1515, gradient/Hessian :1386-1452 / :1516-1583,
//   se3 exponential :1280-1316, WarpPositionList :1338-1348.
// A frame (~300 points, 8.4 KB) stays in L1/L2 across iterations; the 6x6 (3x3) system is reduced
// with warp shuffles and solved redundantly in registers by every lane with Eigen's pivoted LDL^T
// restated with static indexing only.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ba_b200.h"

namespace bapo {

struct Pose {
  float R[9];
  float t[3];
};

__device__ __forceinline__ Pose pose_identity() {
  Pose p;
#pragma unroll
  for (int i = 0; i < 9; ++i) p.R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  p.t[0] = p.t[1] = p.t[2] = 0.0f;
  return p;
}
__device__ __forceinline__ Pose pose_load(const float *g) {
  Pose p;
#pragma unroll
  for (int i = 0; i < 9; ++i) p.R[i] = g[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p.t[i] = g[9 + i];
  return p;
}
__device__ __forceinline__ void pose_store(const Pose &p, float *g) {
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = p.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) g[9 + i] = p.t[i];
}
__device__ __forceinline__ Pose pose_inverse(const Pose &p) {  // Isometry inverse (R^T, -R^T t)
  Pose q;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) q.R[r * 3 + c] = p.R[c * 3 + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) q.t[r] = -(q.R[r * 3] * p.t[0] + q.R[r * 3 + 1] * p.t[1] + q.R[r * 3 + 2] * p.t[2]);
  return q;
}
__device__ __forceinline__ Pose pose_mul(const Pose &a, const Pose &b) {
  Pose q;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      q.R[r * 3 + c] = a.R[r * 3] * b.R[c] + a.R[r * 3 + 1] * b.R[3 + c] + a.R[r * 3 + 2] * b.R[6 + c];
    q.t[r] = a.R[r * 3] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
  return q;
}
__device__ __forceinline__ void pose_apply(const Pose &p, const float *x, float *o) {
#pragma unroll
  for (int r = 0; r < 3; ++r) o[r] = p.R[r * 3] * x[0] + p.R[r * 3 + 1] * x[1] + p.R[r * 3 + 2] * x[2] + p.t[r];
}

// CalculateMatrixExpoenetial_se3<float> (:1280-1316)
__device__ __forceinline__ Pose se3_exp(const float *xi) {
  const float v0 = xi[0], v1 = xi[1], v2 = xi[2], w0 = xi[3], w1 = xi[4], w2 = xi[5];
  const float theta = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
  const float wx[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
  float wx2[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) wx2[r * 3 + c] = wx[r * 3] * wx[c] + wx[r * 3 + 1] * wx[3 + c] + wx[r * 3 + 2] * wx[6 + c];
  float a, b, g;
  if (theta < 1e-7) {
    a = 1.0f; b = 0.5f; g = 0.33333333333333333333333333f;
  } else {
    const float s = sinf(theta), c = cosf(theta);
    a = s / theta;
    b = (1.0f - c) / (theta * theta);
    g = (theta - s) / (theta * theta * theta);
  }
  Pose p;
  float V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float I = (i % 4 == 0) ? 1.0f : 0.0f;
    p.R[i] = I + a * wx[i] + b * wx2[i];
    V[i] = I + b * wx[i] + g * wx2[i];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) p.t[r] = V[r * 3] * v0 + V[r * 3 + 1] * v1 + V[r * 3 + 2] * v2;
  return p;
}

// Eigen::LDLT (diagonal-pivoted, left-looking, unblocked) + solve with the |D| <= FLT_MIN -> 0
// rule, N x N in registers with static indexing only.  H: packed upper (row-major r<=c), already
// damped.  g -> x.
template <int N>
__device__ __forceinline__ void ldlt_solve(const float *Hu, const float *g, float *x) {
  float M[N][N];
  {
    int e = 0;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = r; c < N; ++c) { M[r][c] = Hu[e]; M[c][r] = Hu[e]; ++e; }
  }
  int tr[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int p = k;
    float best = fabsf(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const float v = fabsf(M[i][i]);
      if (v > best) { best = v; p = i; }
    }
    tr[k] = p;
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      if (p == i) {
        // symmetric interchange of indices k and i: rows over all columns, then columns over the
        // trailing rows (columns < k hold L and are only row-swapped)
#pragma unroll
        for (int c = 0; c < N; ++c) { const float t = M[k][c]; M[k][c] = M[i][c]; M[i][c] = t; }
#pragma unroll
        for (int r = k; r < N; ++r) { const float t = M[r][k]; M[r][k] = M[r][i]; M[r][i] = t; }
      }
    }
    float acc = 0.0f;
    float temp[N];
#pragma unroll
    for (int c = 0; c < k; ++c) { temp[c] = M[c][c] * M[k][c]; acc += M[k][c] * temp[c]; }
    if (k > 0) M[k][k] -= acc;
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < k; ++c) s += M[r][c] * temp[c];
      if (k > 0) M[r][k] -= s;
    }
    const float akk = M[k][k];
    if (fabsf(akk) > 0.0f) {
#pragma unroll
      for (int r = k + 1; r < N; ++r) M[r][k] /= akk;
    }
  }
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = g[i];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (tr[k] == i) { const float t = v[k]; v[k] = v[i]; v[i] = t; }
#pragma unroll
  for (int c = 0; c < N; ++c)
#pragma unroll
    for (int r = c + 1; r < N; ++r) v[r] -= M[r][c] * v[c];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = (fabsf(M[i][i]) > 1.17549435e-38f) ? v[i] / M[i][i] : 0.0f;
#pragma unroll
  for (int r = N - 1; r >= 0; --r) {
    float acc = v[r];
#pragma unroll
    for (int i = r + 1; i < N; ++i) acc -= M[i][r] * v[i];
    v[r] = acc;
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k)
#pragma unroll
    for (int i = k + 1; i < N; ++i)
      if (tr[k] == i) { const float t = v[k]; v[k] = v[i]; v[i] = t; }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = v[i];
}

// gradient / Hessian accumulation (:1386-1452, :1516-1583), D = 6 or 3; packed upper H.
template <int D>
__device__ __forceinline__ void grad_hess(const float *Ju, const float *Jv, float ru, float rv, float thres,
                                          float *H, float *g, float &err, float &enw) {
  const float s = fabsf(ru) + fabsf(rv);
  enw = s;
  if (s >= thres) {
    const float w = thres / s;
    const float wru = w * ru, wrv = w * rv;
    int e = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
      const float wJu = w * Ju[r], wJv = w * Jv[r];
#pragma unroll
      for (int c = r; c < D; ++c) H[e++] += (wJu * Ju[c] + wJv * Jv[c]);
    }
#pragma unroll
    for (int r = 0; r < D; ++r) g[r] -= (wru * Ju[r] + wrv * Jv[r]);
    err += wru * ru;  // u term only (:1432)
  } else {
    int e = 0;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) H[e++] += (Ju[r] * Ju[c] + Jv[r] * Jv[c]);
#pragma unroll
    for (int r = 0; r < D; ++r) g[r] -= (ru * Ju[r] + rv * Jv[r]);
    err += rv * rv;   // v term only (:1450)
  }
}

struct Args {
  int kind, n_frames, max_iter;
  const int *offsets;
  const float *points, *pxl, *pxr;
  float intr_l[4], intr_r[4];
  float l2r[12], b2c[12];
  const float *w2l;     // per frame (planar)
  const float *poses_in;
  float *poses_out;
  uint8_t *mask_l, *mask_r;
  ba_poseonly_result *results;
  float *hist_cost, *hist_step, *debug_poses;
  float thr_step, thr_cost, thr_huber, thr_outlier;
};

template <int G>
__device__ __forceinline__ float group_sum(float v, float *sm /*[G/32]*/) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (G > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) s += sm[w];
    v = s;
  }
  return v;
}

// KIND: 0/1 six-dof mono/stereo, 2/3 planar mono/stereo.  G threads per frame (32: warp per frame,
// several frames per CTA; otherwise one CTA per frame).
template <int KIND, int G>
__global__ void __launch_bounds__(G == 32 ? 128 : G) k_poseonly(Args a) {
  constexpr bool STEREO = (KIND & 1) != 0;
  constexpr bool PLANAR = KIND >= 2;
  constexpr int D = PLANAR ? 3 : 6;
  constexpr int NH = D * (D + 1) / 2;
  __shared__ float sm[G > 32 ? G / 32 : 1];
  const int frame = (G == 32) ? (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  if (frame >= a.n_frames) return;
  const int lane = (G == 32) ? (threadIdx.x & 31) : threadIdx.x;
  const int o0 = a.offsets[frame], n_pts = a.offsets[frame + 1] - o0;
  const float *Xw = a.points + 3 * (size_t)o0;
  const float *pl = a.pxl + 2 * (size_t)o0;
  const float *pr = STEREO ? a.pxr + 2 * (size_t)o0 : nullptr;
  uint8_t *ml = a.mask_l ? a.mask_l + o0 : nullptr;
  uint8_t *mr = (STEREO && a.mask_r) ? a.mask_r + o0 : nullptr;
  for (int i = lane; i < n_pts; i += G) {
    if (ml) ml[i] = 1;
    if (mr) mr[i] = 1;
  }
  Pose T_rl = pose_identity();
  if (STEREO) { Pose l2r = pose_load(a.l2r); T_rl = pose_inverse(l2r); }
  const Pose pose_in = pose_load(a.poses_in + 12 * (size_t)frame);
  // six-dof state
  Pose T_cw = pose_inverse(pose_in);
  // planar state
  Pose T_bc = pose_identity(), T_cb = pose_identity(), T_b2b1 = pose_identity(), T_wc_opt = pose_in;
  float R_cb_right[9];
  float prm[3] = {0.f, 0.f, 0.f};
  if (PLANAR) {
    T_bc = pose_load(a.b2c);
    T_cb = pose_inverse(T_bc);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        R_cb_right[r * 3 + c] = T_rl.R[r * 3] * T_cb.R[c] + T_rl.R[r * 3 + 1] * T_cb.R[3 + c] + T_rl.R[r * 3 + 2] * T_cb.R[6 + c];
    const Pose w2l = pose_load(a.w2l + 12 * (size_t)frame);
    const Pose c2c1 = pose_mul(pose_inverse(pose_in), w2l);
    const Pose b2b1 = pose_mul(pose_mul(T_bc, c2c1), T_cb);
    prm[0] = b2b1.t[0]; prm[1] = b2b1.t[1]; prm[2] = atan2f(b2b1.R[3], b2b1.R[0]);
  }
  // number of valid right observations is iteration-invariant (:298)
  float cnt_right = 0.f;
  if (STEREO) {
    for (int i = lane; i < n_pts; i += G) cnt_right += !(pr[2 * i] < 0.f || pr[2 * i + 1] < 0.f) ? 1.f : 0.f;
    cnt_right = group_sum<G>(cnt_right, sm);
  }
  const float inverse_n_pts = 1.0f / (float)n_pts;
  bool is_converged = true;
  float err_prev = 1e10f;
  const float lambda = 1e-5f;
  int n_iter = 0, n_summary = 0;
  float last_err = 0.f, last_step = 0.f;
  for (int iter = 0; iter < a.max_iter; ++iter) {
    float H[NH], g[D];
#pragma unroll
    for (int i = 0; i < NH; ++i) H[i] = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) g[i] = 0.f;
    float err = 0.f;
    float cos_psi = 1.f, sin_psi = 0.f;
    Pose T_left = T_cw, T_right = T_cw;
    if (PLANAR) {
      cos_psi = cosf(prm[2]); sin_psi = sinf(prm[2]);
      T_b2b1 = pose_identity();
      T_b2b1.R[0] = cos_psi; T_b2b1.R[1] = -sin_psi; T_b2b1.R[3] = sin_psi; T_b2b1.R[4] = cos_psi;
      T_b2b1.t[0] = prm[0]; T_b2b1.t[1] = prm[1]; T_b2b1.t[2] = 0.f;
      T_left = pose_mul(T_cb, T_b2b1);
      T_right = pose_mul(T_rl, T_left);
    }
    for (int i = lane; i < n_pts; i += G) {
      const float X[3] = {Xw[3 * i], Xw[3 * i + 1], Xw[3 * i + 2]};
      float Xl[3], Ju[D], Jv[D], ru, rv, enw;
      pose_apply(T_left, X, Xl);
      auto jac = [&](const float *Xc, const float *px, const float *intr, const float *Rcb) {
        const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
        const float inverse_z = 1.0f / Xc[2];
        const float x_inverse_z = Xc[0] * inverse_z, y_inverse_z = Xc[1] * inverse_z;
        const float fx_x_inverse_z = fx * x_inverse_z, fy_y_inverse_z = fy * y_inverse_z;
        ru = (fx_x_inverse_z + cx) - px[0];
        rv = (fy_y_inverse_z + cy) - px[1];
        if (!PLANAR) {
          Ju[0] = fx * inverse_z; Ju[1] = 0.0f; Ju[2] = -fx_x_inverse_z * inverse_z;
          Ju[D > 3 ? 3 : 0] = -fx_x_inverse_z * y_inverse_z;
          Ju[D > 3 ? 4 : 0] = fx * (1.0f + x_inverse_z * x_inverse_z);
          Ju[D > 3 ? 5 : 0] = -fx * y_inverse_z;
          Jv[0] = 0.0f; Jv[1] = fy * inverse_z; Jv[2] = -fy_y_inverse_z * inverse_z;
          Jv[D > 3 ? 3 : 0] = -fy * (1.0f + y_inverse_z * y_inverse_z);
          Jv[D > 3 ? 4 : 0] = fy_y_inverse_z * x_inverse_z;
          Jv[D > 3 ? 5 : 0] = fy * x_inverse_z;
        } else {
          const float r11 = Rcb[0], r12 = Rcb[1], r21 = Rcb[3], r22 = Rcb[4], r31 = Rcb[6], r32 = Rcb[7];
          const float alpha_1 = fx * inverse_z, alpha_2 = -fx_x_inverse_z * inverse_z;
          const float beta_1 = fy * inverse_z, beta_2 = -fy_y_inverse_z * inverse_z;
          const float Aa = -sin_psi * X[0] - cos_psi * X[1];
          const float Bb = cos_psi * X[0] - sin_psi * X[1];
          Ju[0] = alpha_1 * r11 + alpha_2 * r31;
          Ju[1] = alpha_1 * r12 + alpha_2 * r32;
          Ju[2] = Ju[0] * Aa + Ju[1] * Bb;
          Jv[0] = beta_1 * r21 + beta_2 * r31;
          Jv[1] = beta_1 * r22 + beta_2 * r32;
          Jv[2] = Jv[0] * Aa + Jv[1] * Bb;
        }
      };
      jac(Xl, pl + 2 * i, a.intr_l, T_cb.R);
      grad_hess<D>(Ju, Jv, ru, rv, a.thr_huber, H, g, err, enw);
      if (enw >= a.thr_outlier && ml) ml[i] = 0;
      if (STEREO) {
        const float pr0 = pr[2 * i], pr1 = pr[2 * i + 1];
        if (!(pr0 < 0.f || pr1 < 0.f)) {
          float Xr[3];
          if (PLANAR) pose_apply(T_right, X, Xr); else pose_apply(T_rl, Xl, Xr);
          jac(Xr, pr + 2 * i, a.intr_r, R_cb_right);
          grad_hess<D>(Ju, Jv, ru, rv, a.thr_huber, H, g, err, enw);
          if (enw >= a.thr_outlier && mr) mr[i] = 0;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NH; ++i) H[i] = group_sum<G>(H[i], sm);
#pragma unroll
    for (int i = 0; i < D; ++i) g[i] = group_sum<G>(g[i], sm);
    err = group_sum<G>(err, sm);
    // damping (:103 / :333) on the diagonal of the packed upper matrix
    {
      int e = 0;
#pragma unroll
      for (int r = 0; r < D; ++r) { H[e] *= (1.0f + lambda); e += D - r; }
    }
    float delta[D];
    ldlt_solve<D>(H, g, delta);
    if (!PLANAR) {
      const Pose dT = se3_exp(delta);
      T_cw = pose_mul(dT, T_cw);
      if (a.debug_poses && lane == 0) pose_store(pose_inverse(T_cw), a.debug_poses + ((size_t)frame * a.max_iter + iter) * 12);
    } else {
      Pose dT = pose_identity();
      dT.R[0] = cosf(delta[2]); dT.R[1] = -sinf(delta[2]); dT.R[3] = sinf(delta[2]); dT.R[4] = cosf(delta[2]);
      dT.t[0] = delta[0]; dT.t[1] = delta[1]; dT.t[2] = 0.f;
      T_b2b1 = pose_mul(dT, T_b2b1);
      prm[0] = T_b2b1.t[0]; prm[1] = T_b2b1.t[1]; prm[2] += delta[2];
      T_wc_opt = pose_mul(pose_inverse(T_b2b1), T_bc);
      if (a.debug_poses && lane == 0) pose_store(T_wc_opt, a.debug_poses + ((size_t)frame * a.max_iter + iter) * 12);
    }
    if (STEREO) err /= ((float)n_pts + cnt_right) * 0.5f;
    else err *= (inverse_n_pts * 0.5f);
    const float delta_error = fabsf(err - err_prev);
    float nrm = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) nrm += delta[i] * delta[i];
    nrm = sqrtf(nrm);
    n_iter = iter + 1;
    last_err = err; last_step = nrm;
    if (nrm < a.thr_step || delta_error < a.thr_cost) { is_converged = true; break; }
    if (iter == a.max_iter - 1) is_converged = false;
    if (lane == 0) {
      if (a.hist_cost) a.hist_cost[(size_t)frame * a.max_iter + n_summary] = err;
      if (a.hist_step) a.hist_step[(size_t)frame * a.max_iter + n_summary] = nrm;
    }
    ++n_summary;
    err_prev = err;
  }
  if (lane == 0) {
    float nn = 0.f;
    const Pose &chk = PLANAR ? T_b2b1 : T_cw;
#pragma unroll
    for (int i = 0; i < 9; ++i) nn += chk.R[i] * chk.R[i];
    const bool ok = !isnan(sqrtf(nn));
    if (ok) pose_store(PLANAR ? T_wc_opt : pose_inverse(T_cw), a.poses_out + 12 * (size_t)frame);
    else pose_store(pose_in, a.poses_out + 12 * (size_t)frame);
    ba_poseonly_result r;
    r.n_iterations = n_iter; r.converged = is_converged ? 1 : 0; r.success = ok ? 1 : 0;
    r.n_summary = n_summary; r.final_error = last_err; r.final_step = last_step;
    a.results[frame] = r;
  }
}

template <int G>
static void launch_kind(const Args &a, cudaStream_t st) {
  const int grid = (G == 32) ? (a.n_frames + 3) / 4 : a.n_frames;
  const int block = (G == 32) ? 128 : G;
  switch (a.kind) {
    case 0: k_poseonly<0, G><<<grid, block, 0, st>>>(a); break;
    case 1: k_poseonly<1, G><<<grid, block, 0, st>>>(a); break;
    case 2: k_poseonly<2, G><<<grid, block, 0, st>>>(a); break;
    default: k_poseonly<3, G><<<grid, block, 0, st>>>(a); break;
  }
}

}  // namespace bapo

struct ba_poseonly_batch {
  int device = 0, kind = 0, n_frames = 0;
  long long n_pts = 0;
  int *offsets = nullptr;
  float *points = nullptr, *pxl = nullptr, *pxr = nullptr, *w2l = nullptr, *poses_in = nullptr, *poses_out = nullptr;
  uint8_t *mask_l = nullptr, *mask_r = nullptr;
  ba_poseonly_result *results = nullptr;
  float *hist_cost = nullptr, *hist_step = nullptr, *debug_poses = nullptr;
  int hist_iters = 0;
  void *arena = nullptr;       // one pool allocation behind offsets .. results
  float intr_l[4], intr_r[4], l2r[12], b2c[12];
  int group = 32;
};

#define PO_TRY(expr)                                                                  \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      fprintf(stderr, "ba_b200 poseonly: %s: %s\n", #expr, cudaGetErrorString(_e));   \
      return BA_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

extern "C" {

void ba_poseonly_free(ba_poseonly_batch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  // one arena from the device's stream-ordered pool (a dozen cudaMalloc / cudaFree pairs per call cost 5 - 50 ms)
  if (b->arena) cudaFreeAsync(b->arena, 0);
  if (b->hist_cost) cudaFreeAsync(b->hist_cost, 0);
  delete b;
}

int ba_poseonly_upload(ba_poseonly_batch **out, int device, int kind, int n_frames, const int *offsets,
                       const float *points, const float *px_left, const float *px_right,
                       const float *intr_left, const float *intr_right, const float *left_to_right,
                       const float *base_to_camera, const float *world_to_last, const float *poses_init) {
  if (!out || kind < 0 || kind > 3 || n_frames < 0 || !offsets || !points || !px_left || !intr_left || !poses_init)
    return BA_ERR_INVALID;
  const bool stereo = kind & 1, planar = kind >= 2;
  if (stereo && (!px_right || !left_to_right)) return BA_ERR_INVALID;
  if (planar && (!base_to_camera || !world_to_last)) return BA_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    fprintf(stderr, "ba_b200: no CUDA device available; the pose-only engine has no CPU fallback\n");
    return BA_ERR_CUDA;
  }
  PO_TRY(cudaSetDevice(device));
  ba_poseonly_batch *b = new ba_poseonly_batch();
  b->device = device; b->kind = kind; b->n_frames = n_frames;
  const long long n = n_frames > 0 ? offsets[n_frames] : 0;
  b->n_pts = n;
  std::memcpy(b->intr_l, intr_left, 16);
  std::memcpy(b->intr_r, intr_right ? intr_right : intr_left, 16);
  const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  std::memcpy(b->l2r, left_to_right ? left_to_right : ident, 48);
  std::memcpy(b->b2c, base_to_camera ? base_to_camera : ident, 48);
  const size_t nz = (size_t)std::max<long long>(n, 1), fz = (size_t)std::max(n_frames, 1);
  {
    static bool pool_ready[64] = {};
    if (device >= 0 && device < 64 && !pool_ready[device]) {   // keep freed blocks in the pool between calls
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready[device] = true;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t sz[10] = {al((fz + 1) * sizeof(int)), al(nz * 12), al(nz * 8), al(nz * 8), al(fz * 48), al(fz * 48), al(fz * 48),
                           al(nz), al(nz), al(fz * sizeof(ba_poseonly_result))};
    size_t total = 0;
    for (size_t v : sz) total += v;
    PO_TRY(cudaMallocAsync(&b->arena, total, 0));
    char *p = static_cast<char *>(b->arena);
    b->offsets = reinterpret_cast<decltype(b->offsets)>(p); p += sz[0];
    b->points = reinterpret_cast<decltype(b->points)>(p); p += sz[1];
    b->pxl = reinterpret_cast<decltype(b->pxl)>(p); p += sz[2];
    b->pxr = reinterpret_cast<decltype(b->pxr)>(p); p += sz[3];
    b->w2l = reinterpret_cast<decltype(b->w2l)>(p); p += sz[4];
    b->poses_in = reinterpret_cast<decltype(b->poses_in)>(p); p += sz[5];
    b->poses_out = reinterpret_cast<decltype(b->poses_out)>(p); p += sz[6];
    b->mask_l = reinterpret_cast<decltype(b->mask_l)>(p); p += sz[7];
    b->mask_r = reinterpret_cast<decltype(b->mask_r)>(p); p += sz[8];
    b->results = reinterpret_cast<decltype(b->results)>(p);
  }
  PO_TRY(cudaMemcpy(b->offsets, offsets, ((size_t)n_frames + 1) * sizeof(int), cudaMemcpyHostToDevice));
  PO_TRY(cudaMemcpy(b->points, points, (size_t)n * 12, cudaMemcpyHostToDevice));
  PO_TRY(cudaMemcpy(b->pxl, px_left, (size_t)n * 8, cudaMemcpyHostToDevice));
  if (stereo) PO_TRY(cudaMemcpy(b->pxr, px_right, (size_t)n * 8, cudaMemcpyHostToDevice));
  if (planar) PO_TRY(cudaMemcpy(b->w2l, world_to_last, (size_t)n_frames * 48, cudaMemcpyHostToDevice));
  PO_TRY(cudaMemcpy(b->poses_in, poses_init, (size_t)n_frames * 48, cudaMemcpyHostToDevice));
  // one warp per frame for VO-sized frames, one CTA per frame for large ones
  const long long avg = n_frames > 0 ? n / n_frames : 0;
  b->group = avg > 2048 ? 256 : 32;
  *out = b;
  return BA_OK;
}

static int po_alloc_hist(ba_poseonly_batch *b, int max_iter) {
  if (b->hist_iters >= max_iter && b->hist_cost) return BA_OK;
  if (b->hist_cost) cudaFreeAsync(b->hist_cost, 0);
  const size_t fz = (size_t)std::max(b->n_frames, 1) * std::max(max_iter, 1);
  const size_t a4 = (fz * 4 + 255) / 256 * 256;
  void *blk = nullptr;
  PO_TRY(cudaMallocAsync(&blk, 2 * a4 + fz * 48, 0));       // hist_cost | hist_step | debug_poses
  b->hist_cost = reinterpret_cast<decltype(b->hist_cost)>(blk);
  b->hist_step = reinterpret_cast<decltype(b->hist_step)>(static_cast<char *>(blk) + a4);
  b->debug_poses = reinterpret_cast<decltype(b->debug_poses)>(static_cast<char *>(blk) + 2 * a4);
  b->hist_iters = max_iter;
  return BA_OK;
}

static int po_run(ba_poseonly_batch *b, const ba_poseonly_options *opt, cudaStream_t st, bool want_hist) {
  bapo::Args a;
  a.kind = b->kind; a.n_frames = b->n_frames; a.max_iter = opt->max_num_iterations;
  a.offsets = b->offsets; a.points = b->points; a.pxl = b->pxl; a.pxr = b->pxr;
  std::memcpy(a.intr_l, b->intr_l, 16); std::memcpy(a.intr_r, b->intr_r, 16);
  std::memcpy(a.l2r, b->l2r, 48); std::memcpy(a.b2c, b->b2c, 48);
  a.w2l = b->w2l; a.poses_in = b->poses_in; a.poses_out = b->poses_out;
  a.mask_l = b->mask_l; a.mask_r = b->mask_r; a.results = b->results;
  a.hist_cost = want_hist ? b->hist_cost : nullptr;
  a.hist_step = want_hist ? b->hist_step : nullptr;
  a.debug_poses = want_hist ? b->debug_poses : nullptr;
  a.thr_step = opt->threshold_step_size; a.thr_cost = opt->threshold_cost_change;
  a.thr_huber = opt->threshold_huber_loss; a.thr_outlier = opt->threshold_outlier_rejection;
  if (b->n_frames == 0) return BA_OK;
  if (b->group == 32) bapo::launch_kind<32>(a, st); else bapo::launch_kind<256>(a, st);
  PO_TRY(cudaGetLastError());
  return BA_OK;
}

int ba_poseonly_run(ba_poseonly_batch *b, const ba_poseonly_options *opt, void *cuda_stream) {
  if (!b || !opt) return BA_ERR_INVALID;
  PO_TRY(cudaSetDevice(b->device));
  return po_run(b, opt, (cudaStream_t)cuda_stream, false);
}

int ba_poseonly_download(ba_poseonly_batch *b, float *poses_out, uint8_t *mask_left, uint8_t *mask_right,
                         ba_poseonly_result *results) {
  if (!b) return BA_ERR_INVALID;
  PO_TRY(cudaSetDevice(b->device));
  PO_TRY(cudaDeviceSynchronize());
  if (poses_out) PO_TRY(cudaMemcpy(poses_out, b->poses_out, (size_t)b->n_frames * 48, cudaMemcpyDeviceToHost));
  if (mask_left) PO_TRY(cudaMemcpy(mask_left, b->mask_l, (size_t)b->n_pts, cudaMemcpyDeviceToHost));
  if (mask_right && (b->kind & 1)) PO_TRY(cudaMemcpy(mask_right, b->mask_r, (size_t)b->n_pts, cudaMemcpyDeviceToHost));
  if (results) PO_TRY(cudaMemcpy(results, b->results, (size_t)b->n_frames * sizeof(ba_poseonly_result), cudaMemcpyDeviceToHost));
  return BA_OK;
}

int ba_poseonly_solve_batched(int device, int kind, int n_frames, const int *offsets, const float *points,
                              const float *px_left, const float *px_right, const float *intr_left,
                              const float *intr_right, const float *left_to_right,
                              const float *base_to_camera, const float *world_to_last, float *poses_io,
                              uint8_t *mask_left, uint8_t *mask_right, const ba_poseonly_options *opt,
                              ba_poseonly_result *results, float *hist_cost, float *hist_step,
                              float *debug_poses) {
  if (!opt || !poses_io) return BA_ERR_INVALID;
  ba_poseonly_batch *b = nullptr;
  int rc = ba_poseonly_upload(&b, device, kind, n_frames, offsets, points, px_left, px_right, intr_left,
                              intr_right, left_to_right, base_to_camera, world_to_last, poses_io);
  if (rc) return rc;
  const bool want_hist = hist_cost || hist_step || debug_poses;
  if (want_hist) { rc = po_alloc_hist(b, opt->max_num_iterations); if (rc) { ba_poseonly_free(b); return rc; } }
  rc = po_run(b, opt, nullptr, want_hist);
  if (!rc) rc = ba_poseonly_download(b, poses_io, mask_left, mask_right, results);
  if (!rc && want_hist) {
    const size_t fz = (size_t)n_frames * std::max(opt->max_num_iterations, 0);
    if (hist_cost) cudaMemcpy(hist_cost, b->hist_cost, fz * 4, cudaMemcpyDeviceToHost);
    if (hist_step) cudaMemcpy(hist_step, b->hist_step, fz * 4, cudaMemcpyDeviceToHost);
    if (debug_poses) cudaMemcpy(debug_poses, b->debug_poses, fz * 48, cudaMemcpyDeviceToHost);
  }
  ba_poseonly_free(b);
  return rc;
}

}  // extern "C"
