// geometry_library.h -- drop-in for the reference's utility/geometry_library.h (namespace geometry, same function
// names and argument order, Eigen-typed; /root/reference/utility/geometry_library.h:9-54).  The arithmetic lives in
// geometry_math.h, which the CUDA kernels share (ba_geometry_batched runs the same functions one element per
// thread); these wrappers only convert between Eigen's column-major storage and the row-major arrays used there.
#ifndef BA_B200_GEOMETRY_LIBRARY_H_
#define BA_B200_GEOMETRY_LIBRARY_H_

#include <cmath>
#include <iostream>

#include "../eigen_shim.h"
#include "geometry_math.h"

namespace geometry {
namespace detail {
template <typename T> inline void to_rm(const Eigen::Matrix<T, 3, 3> &R, T *o) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = R(r, c); }
template <typename T> inline Eigen::Matrix<T, 3, 3> from_rm(const T *o) { Eigen::Matrix<T, 3, 3> R; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = o[3 * r + c]; return R; }
template <typename T> inline void split(const Eigen::Matrix<T, 4, 4> &M, T *R, T *t) { for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = M(r, c); t[r] = M(r, 3); } }
template <typename T> inline Eigen::Matrix<T, 4, 4> join(const T *R, const T *t) {
  Eigen::Matrix<T, 4, 4> M;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M(r, c) = T(0);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) M(r, c) = R[3 * r + c]; M(r, 3) = t[r]; }
  M(3, 3) = T(1);
  return M;
}
template <typename T> inline Eigen::Matrix<T, 3, 3> skew(const Eigen::Matrix<T, 3, 1> &v) {   // :6-21
  Eigen::Matrix<T, 3, 3> S;
  S(0, 0) = 0; S(0, 1) = -v(2); S(0, 2) = v(1);
  S(1, 0) = v(2); S(1, 1) = 0; S(1, 2) = -v(0);
  S(2, 0) = -v(1); S(2, 1) = v(0); S(2, 2) = 0;
  return S;
}
template <typename T> inline Eigen::Matrix<T, 4, 4> q_mult_matrix(const Eigen::Matrix<T, 4, 1> &q, T s) {   // :23-59, s = -1 right, +1 left
  Eigen::Matrix<T, 4, 4> O;
  O(0, 0) = q(0); O(0, 1) = -q(1); O(0, 2) = -q(2); O(0, 3) = -q(3);
  O(1, 0) = q(1); O(1, 1) = q(0); O(1, 2) = -s * q(3); O(1, 3) = s * q(2);
  O(2, 0) = q(2); O(2, 1) = s * q(3); O(2, 2) = q(0); O(2, 3) = -s * q(1);
  O(3, 0) = q(3); O(3, 1) = -s * q(2); O(3, 2) = s * q(1); O(3, 3) = q(0);
  return O;
}
template <typename T, int N> inline Eigen::Matrix<T, N, 1> vec(const T *o) { Eigen::Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v(i) = o[i]; return v; }
template <typename T> inline Eigen::Matrix<T, 3, 3> q2r_t(const Eigen::Matrix<T, 4, 1> &q) { T R[9]; const T qq[4] = {q(0), q(1), q(2), q(3)}; ba_geom::q2r(qq, R); return from_rm(R); }
template <typename T> inline Eigen::Matrix<T, 4, 1> r2q_t(const Eigen::Matrix<T, 3, 3> &Rm) { T R[9], q[4]; to_rm(Rm, R); ba_geom::r2q(R, q); return vec<T, 4>(q); }
template <typename T> inline void se3Exp_t(const Eigen::Matrix<T, 6, 1> &xi, Eigen::Matrix<T, 4, 4> &Tm) { T x[6], R[9], t[3]; for (int i = 0; i < 6; ++i) x[i] = xi(i); ba_geom::se3_exp(x, R, t); Tm = join(R, t); }
template <typename T> inline void SE3Log_t(const Eigen::Matrix<T, 4, 4> &Tm, Eigen::Matrix<T, 6, 1> &xi) {
  T R[9], t[3], x[6];
  split(Tm, R, t);
  ba_geom::se3_log(R, t, x);
  for (int i = 0; i < 6; ++i) xi(i) = x[i];
  if (std::isnan(xi.norm())) std::cout << "============ SE3Log NAN!! ============\n";   // :524-541
}
}  // namespace detail

typedef Eigen::Matrix<double, 4, 1> Vector4d_;
typedef Eigen::Matrix<float, 4, 1> Vector4f_;
typedef Eigen::Matrix<double, 6, 1> Vector6d_;
typedef Eigen::Matrix<float, 6, 1> Vector6f_;

inline Eigen::Matrix3d skewMat(const Eigen::Vector3d &v) { return detail::skew<double>(v); }
inline Eigen::Matrix3f skewMat_f(const Eigen::Vector3f &v) { return detail::skew<float>(v); }
inline Eigen::Matrix4d q_right_mult(const Vector4d_ &q) { return detail::q_mult_matrix<double>(q, -1.0); }
inline Eigen::Matrix4f q_right_mult_f(const Vector4f_ &q) { return detail::q_mult_matrix<float>(q, -1.0f); }
inline Eigen::Matrix4d q_left_mult(const Vector4d_ &q) { return detail::q_mult_matrix<double>(q, 1.0); }
inline Eigen::Matrix4f q_left_mult_f(const Vector4f_ &q) { return detail::q_mult_matrix<float>(q, 1.0f); }
inline Vector4d_ q_conj(const Vector4d_ &q) { Vector4d_ o; o(0) = q(0); o(1) = -q(1); o(2) = -q(2); o(3) = -q(3); return o; }
inline Vector4f_ q_conj_f(const Vector4f_ &q) { Vector4f_ o; o(0) = q(0); o(1) = -q(1); o(2) = -q(2); o(3) = -q(3); return o; }
inline Vector4d_ q1_mult_q2(const Vector4d_ &a, const Vector4d_ &b) { const double x[4] = {a(0), a(1), a(2), a(3)}, y[4] = {b(0), b(1), b(2), b(3)}; double o[4]; ba_geom::q_mult(x, y, o); return detail::vec<double, 4>(o); }
inline Vector4f_ q1_mult_q2_f(const Vector4f_ &a, const Vector4f_ &b) { const float x[4] = {a(0), a(1), a(2), a(3)}, y[4] = {b(0), b(1), b(2), b(3)}; float o[4]; ba_geom::q_mult(x, y, o); return detail::vec<float, 4>(o); }
inline Eigen::Matrix3d q2r(const Vector4d_ &q) { return detail::q2r_t<double>(q); }
inline Eigen::Matrix3f q2r_f(const Vector4f_ &q) { return detail::q2r_t<float>(q); }
inline Vector4d_ rotvec2q(const Eigen::Vector3d &w) { const double x[3] = {w(0), w(1), w(2)}; double q[4]; ba_geom::rotvec2q(x, q); return detail::vec<double, 4>(q); }
inline Vector4f_ rotvec2q_f(const Eigen::Vector3f &w) { const float x[3] = {w(0), w(1), w(2)}; float q[4]; ba_geom::rotvec2q(x, q); return detail::vec<float, 4>(q); }
inline Eigen::Matrix3d a2r(double r, double p, double y) { const double a[3] = {r, p, y}; double R[9]; ba_geom::a2r(a, R); return detail::from_rm(R); }
inline Eigen::Matrix3f a2r_f(float r, float p, float y) { const float a[3] = {r, p, y}; float R[9]; ba_geom::a2r(a, R); return detail::from_rm(R); }
inline Vector4d_ r2q(const Eigen::Matrix3d &R) { return detail::r2q_t<double>(R); }
inline Vector4f_ r2q_f(const Eigen::Matrix3f &R) { return detail::r2q_t<float>(R); }
inline Eigen::Vector3d r2euler(const Eigen::Matrix3d &Rm) { double R[9], e[3]; detail::to_rm(Rm, R); ba_geom::r2euler(R, e); return detail::vec<double, 3>(e); }
inline Eigen::Vector3f r2euler_f(const Eigen::Matrix3f &Rm) { float R[9], e[3]; detail::to_rm(Rm, R); ba_geom::r2euler(R, e); return detail::vec<float, 3>(e); }

inline void se3Exp(const Vector6d_ &xi, Eigen::Matrix4d &T) { detail::se3Exp_t<double>(xi, T); }
inline void se3Exp_f(const Vector6f_ &xi, Eigen::Matrix4f &T) { detail::se3Exp_t<float>(xi, T); }
inline void SE3Log(const Eigen::Matrix4d &T, Vector6d_ &xi) { detail::SE3Log_t<double>(T, xi); }
inline void SE3Log_f(const Eigen::Matrix4f &T, Vector6f_ &xi) { detail::SE3Log_t<float>(T, xi); }
inline void so3Exp(const Eigen::Vector3d &w, Eigen::Matrix3d &R) { const double x[3] = {w(0), w(1), w(2)}; double o[9]; ba_geom::so3_exp(x, o); R = detail::from_rm(o); }
inline void so3Exp(const double w1, const double w2, const double w3, Eigen::Matrix3d &R) { const double x[3] = {w1, w2, w3}; double o[9]; ba_geom::so3_exp(x, o); R = detail::from_rm(o); }
inline void so3Exp_f(const Eigen::Vector3f &w, Eigen::Matrix3f &R) { const float x[3] = {w(0), w(1), w(2)}; float o[9]; ba_geom::so3_exp(x, o); R = detail::from_rm(o); }
// NB the reference takes `w` by reference and fills it (:659, :681)
inline void SO3Log(const Eigen::Matrix3d &Rm, Eigen::Vector3d &w) { double R[9], o[3]; detail::to_rm(Rm, R); ba_geom::so3_log(R, o); w = detail::vec<double, 3>(o); }
inline void SO3Log_f(const Eigen::Matrix3f &Rm, Eigen::Vector3f &w) { float R[9], o[3]; detail::to_rm(Rm, R); ba_geom::so3_log(R, o); w = detail::vec<float, 3>(o); }
inline void addFrontse3(Vector6d_ &xi, const Vector6d_ &dxi) { double a[6], b[6], o[6]; for (int i = 0; i < 6; ++i) { a[i] = xi(i); b[i] = dxi(i); } ba_geom::add_front_se3(a, b, o); for (int i = 0; i < 6; ++i) xi(i) = o[i]; }
inline void addFrontse3_f(Vector6f_ &xi, const Vector6f_ &dxi) { float a[6], b[6], o[6]; for (int i = 0; i < 6; ++i) { a[i] = xi(i); b[i] = dxi(i); } ba_geom::add_front_se3(a, b, o); for (int i = 0; i < 6; ++i) xi(i) = o[i]; }
inline Eigen::Matrix4d inverseSE3(const Eigen::Matrix4d &T) { double R[9], t[3], Ri[9], ti[3]; detail::split(T, R, t); ba_geom::inverse_se3(R, t, Ri, ti); return detail::join(Ri, ti); }
inline Eigen::Matrix4f inverseSE3_f(const Eigen::Matrix4f &T) { float R[9], t[3], Ri[9], ti[3]; detail::split(T, R, t); ba_geom::inverse_se3(R, t, Ri, ti); return detail::join(Ri, ti); }
}  // namespace geometry

#endif  // BA_B200_GEOMETRY_LIBRARY_H_
