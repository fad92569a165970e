// geometry_math.h -- SE(3) / SO(3) / quaternion / Euler helpers of the reference's utility/geometry_library.cpp as
// scalar functions on plain arrays, usable on the host AND inside CUDA kernels (BA_HD).  One source for the C++
// drop-in header (ba_b200/utility/geometry_library.h) and the batched device entry point ba_geometry_batched
// (csrc/ba_geometry.cuh).  Conventions: rotation matrices row-major R[9]; quaternions (w, x, y, z); twists
// xi = [v; w] (translation first), like the reference.  Line numbers: /root/reference/utility/geometry_library.cpp.
#ifndef BA_B200_GEOMETRY_MATH_H_
#define BA_B200_GEOMETRY_MATH_H_

#include <cmath>

#if defined(__CUDACC__)
#define BA_HD __host__ __device__ __forceinline__
#else
#define BA_HD inline
#endif

namespace ba_geom {

template <typename T> BA_HD T t_sqrt(T x) { return sqrt(x); }
template <typename T> BA_HD T t_sin(T x) { return sin(x); }
template <typename T> BA_HD T t_cos(T x) { return cos(x); }
template <typename T> BA_HD T t_acos(T x) { return acos(x); }
template <typename T> BA_HD T t_atan2(T y, T x) { return atan2(y, x); }

// R = I + a [w]x + b [w]x^2   ([w]x^2 = w w^T - |w|^2 I)
template <typename T>
BA_HD void rodrigues(const T *w, T a, T b, T *R) {
  const T w0 = w[0], w1 = w[1], w2 = w[2];
  const T xx = w0 * w0, yy = w1 * w1, zz = w2 * w2;
  R[0] = T(1) - b * (yy + zz); R[1] = -a * w2 + b * w0 * w1;  R[2] = a * w1 + b * w0 * w2;
  R[3] = a * w2 + b * w0 * w1; R[4] = T(1) - b * (xx + zz);   R[5] = -a * w0 + b * w1 * w2;
  R[6] = -a * w1 + b * w0 * w2; R[7] = a * w0 + b * w1 * w2;  R[8] = T(1) - b * (xx + yy);
}

// so3Exp (:590-611): theta < 1e-9 -> I + [w]x + 0.5 [w]x^2, else Rodrigues
template <typename T>
BA_HD void so3_exp(const T *w, T *R) {
  const T th = t_sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < T(1e-9)) {
    rodrigues(w, T(1), T(0.5), R);
  } else {
    const T inv2 = T(1) / (th * th);
    rodrigues(w, t_sin(th) / th, (T(1) - t_cos(th)) * inv2, R);
  }
}

// SO3Log (:659-679): (tr R - 1) / 2 >= 0.999999999 -> 0, else w = theta / (2 sin theta) vee(R - R^T)
template <typename T>
BA_HD void so3_log(const T *R, T *w) {
  const T in_cos = (R[0] + R[4] + R[8] - T(1)) * T(0.5);
  if (in_cos >= T(0.999999999)) {
    w[0] = w[1] = w[2] = T(0);
    return;
  }
  const T th = t_acos(in_cos);
  const T k = th / (T(2) * t_sin(th));
  // lnR = k (R - R^T);  w = (-lnR(1,2), lnR(0,2), -lnR(0,1))
  w[0] = -k * (R[5] - R[7]);
  w[1] = k * (R[2] - R[6]);
  w[2] = -k * (R[1] - R[3]);
}

// se3Exp (:370-427): xi = [v; w]; R as so3Exp, t = V v with V = I + b [w]x + c [w]x^2
template <typename T>
BA_HD void se3_exp(const T *xi, T *R, T *t) {
  const T *v = xi, *w = xi + 3;
  const T th = t_sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  T V[9];
  if (th < T(1e-9)) {
    rodrigues(w, T(1), T(0.5), R);
    rodrigues(w, T(0.5), T(0.33333333333333333333333333), V);
  } else {
    const T inv2 = T(1) / (th * th);
    rodrigues(w, t_sin(th) / th, (T(1) - t_cos(th)) * inv2, R);
    rodrigues(w, (T(1) - t_cos(th)) * inv2, (th - t_sin(th)) / (th * th * th), V);
  }
  for (int r = 0; r < 3; ++r) t[r] = V[3 * r] * v[0] + V[3 * r + 1] * v[1] + V[3 * r + 2] * v[2];
}

// SE3Log (:488-546): w as SO3Log; v = Vin t, Vin = I - 0.5 [w]x + (1 - A / (2 B)) / theta^2 [w]x^2,
// A = sin(theta) / theta, B = (1 - cos(theta)) / theta^2
template <typename T>
BA_HD void se3_log(const T *R, const T *t, T *xi) {
  const T in_cos = (R[0] + R[4] + R[8] - T(1)) * T(0.5);
  T w[3] = {T(0), T(0), T(0)};
  T Vin[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
  if (!(in_cos >= T(0.999999999))) {
    const T th = t_acos(in_cos);
    const T inv = T(1) / th, inv2 = inv * inv;
    const T k = th / (T(2) * t_sin(th));
    w[0] = -k * (R[5] - R[7]);
    w[1] = k * (R[2] - R[6]);
    w[2] = -k * (R[1] - R[3]);
    const T A = t_sin(th) * inv, B = (T(1) - t_cos(th)) * inv2;
    rodrigues(w, T(-0.5), inv2 * (T(1) - A / (T(2) * B)), Vin);
  }
  for (int r = 0; r < 3; ++r) xi[r] = Vin[3 * r] * t[0] + Vin[3 * r + 1] * t[1] + Vin[3 * r + 2] * t[2];
  xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
}

// inverseSE3 (:721-736)
template <typename T>
BA_HD void inverse_se3(const T *R, const T *t, T *Ri, T *ti) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ri[3 * r + c] = R[3 * c + r];
  for (int r = 0; r < 3; ++r) ti[r] = -(Ri[3 * r] * t[0] + Ri[3 * r + 1] * t[1] + Ri[3 * r + 2] * t[2]);
}

// addFrontse3 (:703-719): xi <- Log(Exp(dxi) Exp(xi))
template <typename T>
BA_HD void add_front_se3(const T *xi, const T *dxi, T *out) {
  T R[9], t[3], dR[9], dt[3], Rn[9], tn[3];
  se3_exp(xi, R, t);
  se3_exp(dxi, dR, dt);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rn[3 * r + c] = dR[3 * r] * R[c] + dR[3 * r + 1] * R[3 + c] + dR[3 * r + 2] * R[6 + c];
    tn[r] = dR[3 * r] * t[0] + dR[3 * r + 1] * t[1] + dR[3 * r + 2] * t[2] + dt[r];
  }
  se3_log(Rn, tn, out);
}

// q2r (:93-117)
template <typename T>
BA_HD void q2r(const T *q, T *R) {
  const T qw = q[0], qx = q[1], qy = q[2], qz = q[3];
  const T qw2 = qw * qw, qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz;
  const T qxqy = qx * qy, qwqz = qw * qz, qxqz = qx * qz, qwqy = qw * qy, qwqx = qw * qx, qyqz = qy * qz;
  R[0] = qw2 + qx2 - qy2 - qz2; R[1] = T(2) * (qxqy - qwqz);   R[2] = T(2) * (qxqz + qwqy);
  R[3] = T(2) * (qxqy + qwqz);  R[4] = qw2 - qx2 + qy2 - qz2;  R[5] = T(2) * (qyqz - qwqx);
  R[6] = T(2) * (qxqz - qwqy);  R[7] = T(2) * (qyqz + qwqx);   R[8] = qw2 - qx2 - qy2 + qz2;
}

// r2q (:206-262): largest-component branch selection (the reference compares with the bitwise & of two bools)
template <typename T>
BA_HD void r2q(const T *R, T *q) {
  const T m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
  const T tr = m00 + m11 + m22;
  if (tr > T(0)) {
    const T S = t_sqrt(tr + T(1)) * T(2);
    q[0] = T(0.25) * S; q[1] = (m21 - m12) / S; q[2] = (m02 - m20) / S; q[3] = (m10 - m01) / S;
  } else if ((m00 > m11) & (m00 > m22)) {
    const T S = t_sqrt(T(1) + m00 - m11 - m22) * T(2);
    q[0] = (m21 - m12) / S; q[1] = T(0.25) * S; q[2] = (m01 + m10) / S; q[3] = (m02 + m20) / S;
  } else if (m11 > m22) {
    const T S = t_sqrt(T(1) + m11 - m00 - m22) * T(2);
    q[0] = (m02 - m20) / S; q[1] = (m01 + m10) / S; q[2] = T(0.25) * S; q[3] = (m12 + m21) / S;
  } else {
    const T S = t_sqrt(T(1) + m22 - m00 - m11) * T(2);
    q[0] = (m10 - m01) / S; q[1] = (m02 + m20) / S; q[2] = (m12 + m21) / S; q[3] = T(0.25) * S;
  }
}

// rotvec2q (:146-161): theta < 1e-7 -> identity, else (cos(theta/2), w sin(theta/2) / theta) normalised
template <typename T>
BA_HD void rotvec2q(const T *w, T *q) {
  const T th = t_sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < T(1e-7)) {
    q[0] = T(1); q[1] = q[2] = q[3] = T(0);
    return;
  }
  const T s = t_sin(th * T(0.5)) / th;
  q[0] = t_cos(th * T(0.5)); q[1] = w[0] * s; q[2] = w[1] * s; q[3] = w[2] * s;
  const T nrm = t_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= nrm;
}

// q_conj (:61-66), q1_mult_q2 (:74-81: q = q_left_mult(q1) q2, the Hamilton product)
template <typename T>
BA_HD void q_conj(const T *q, T *o) { o[0] = q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = -q[3]; }
template <typename T>
BA_HD void q_mult(const T *a, const T *b, T *o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[1] * b[0] + a[0] * b[1] - a[3] * b[2] + a[2] * b[3];
  o[2] = a[2] * b[0] + a[3] * b[1] + a[0] * b[2] - a[1] * b[3];
  o[3] = a[3] * b[0] - a[2] * b[1] + a[1] * b[2] + a[0] * b[3];
}

// a2r (:181-191): R = Rz(y) Ry(p) Rx(r)
template <typename T>
BA_HD void a2r(const T *rpy, T *R) {
  const T cr = t_cos(rpy[0]), sr = t_sin(rpy[0]), cp = t_cos(rpy[1]), sp = t_sin(rpy[1]), cy = t_cos(rpy[2]), sy = t_sin(rpy[2]);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// r2euler (:322-344)
template <typename T>
BA_HD void r2euler(const T *R, T *e) {
  const T sy = t_sqrt(R[0] * R[0] + R[3] * R[3]);
  if (sy < T(1e-6)) {
    e[0] = t_atan2(-R[5], R[4]); e[1] = t_atan2(-R[6], sy); e[2] = T(0);
  } else {
    e[0] = t_atan2(R[7], R[8]); e[1] = t_atan2(-R[6], sy); e[2] = t_atan2(R[3], R[0]);
  }
}

}  // namespace ba_geom
#endif  // BA_B200_GEOMETRY_MATH_H_
