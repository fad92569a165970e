// eigen_shim.h -- the reference's public API is Eigen-typed (core/full_bundle_adjustment_solver.h:34-67).
// With real Eigen on the include path it is used as is; otherwise (this repo's build container has no
// Eigen) a minimal, layout-compatible subset is provided: fixed-size column-major Matrix and an
// Isometry Transform stored as a column-major 4x4 (128 B for double), exactly Eigen's layouts, with the
// handful of members the reference's callers use (test/test_ba.cpp, README.md:14-69).
#ifndef BA_B200_EIGEN_SHIM_H_
#define BA_B200_EIGEN_SHIM_H_

#if defined(__has_include)
#if __has_include(<eigen3/Eigen/Dense>) && !defined(BA_B200_FORCE_EIGEN_SHIM)
#define BA_B200_HAVE_EIGEN 1
#endif
#endif

#ifdef BA_B200_HAVE_EIGEN
#include <eigen3/Eigen/Dense>
#include <eigen3/Eigen/Geometry>
#else
#include <cmath>
#include <cstddef>
#include <initializer_list>

namespace Eigen {

enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 3, Projective = 4 };

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  Matrix() { for (int i = 0; i < R * C; ++i) m_[i] = T(0); }
  Matrix(T a, T b) { static_assert(R * C == 2, "2 coefficients"); m_[0] = a; m_[1] = b; }
  Matrix(T a, T b, T c) { static_assert(R * C == 3, "3 coefficients"); m_[0] = a; m_[1] = b; m_[2] = c; }
  Matrix(std::initializer_list<T> l) { int i = 0; for (T v : l) if (i < R * C) m_[i++] = v; for (; i < R * C; ++i) m_[i] = T(0); }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() { Matrix I; for (int i = 0; i < (R < C ? R : C); ++i) I(i, i) = T(1); return I; }
  static Matrix UnitX() { Matrix v; v(0) = T(1); return v; }
  static Matrix UnitY() { Matrix v; v(1) = T(1); return v; }
  static Matrix UnitZ() { Matrix v; v(2) = T(1); return v; }
  T &operator()(int r, int c) { return m_[c * R + r]; }
  const T &operator()(int r, int c) const { return m_[c * R + r]; }
  T &operator()(int i) { return m_[i]; }
  const T &operator()(int i) const { return m_[i]; }
  T &operator[](int i) { return m_[i]; }
  const T &operator[](int i) const { return m_[i]; }
  T &x() { return m_[0]; } const T &x() const { return m_[0]; }
  T &y() { return m_[1]; } const T &y() const { return m_[1]; }
  T &z() { return m_[2]; } const T &z() const { return m_[2]; }
  T *data() { return m_; } const T *data() const { return m_; }
  void setZero() { for (int i = 0; i < R * C; ++i) m_[i] = T(0); }
  void setIdentity() { *this = Identity(); }
  T norm() const { T s = 0; for (int i = 0; i < R * C; ++i) s += m_[i] * m_[i]; return std::sqrt(s); }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
  Matrix operator+(const Matrix &o) const { Matrix q; for (int i = 0; i < R * C; ++i) q.m_[i] = m_[i] + o.m_[i]; return q; }
  Matrix operator-(const Matrix &o) const { Matrix q; for (int i = 0; i < R * C; ++i) q.m_[i] = m_[i] - o.m_[i]; return q; }
  Matrix operator-() const { Matrix q; for (int i = 0; i < R * C; ++i) q.m_[i] = -m_[i]; return q; }
  Matrix operator*(T s) const { Matrix q; for (int i = 0; i < R * C; ++i) q.m_[i] = m_[i] * s; return q; }
  Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) m_[i] += o.m_[i]; return *this; }
  Matrix &operator*=(T s) { for (int i = 0; i < R * C; ++i) m_[i] *= s; return *this; }
  template <int K>
  Matrix<T, R, K> operator*(const Matrix<T, C, K> &o) const {
    Matrix<T, R, K> q;
    for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) { T s = 0; for (int c = 0; c < C; ++c) s += (*this)(r, c) * o(c, k); q(r, k) = s; }
    return q;
  }
 private:
  T m_[R * C];
};
template <typename T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C> &m) { return m * s; }

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 4, 4> Matrix4d; typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 1> Vector4d; typedef Matrix<float, 4, 1> Vector4f;

template <typename T>
class AngleAxis {
 public:
  AngleAxis(T angle, const Matrix<T, 3, 1> &axis) : a_(angle), k_(axis) {}
  Matrix<T, 3, 3> toRotationMatrix() const {
    const T c = std::cos(a_), s = std::sin(a_), v = T(1) - c;
    const T x = k_(0), y = k_(1), z = k_(2);
    Matrix<T, 3, 3> R;
    R(0, 0) = c + x * x * v;     R(0, 1) = x * y * v - z * s; R(0, 2) = x * z * v + y * s;
    R(1, 0) = y * x * v + z * s; R(1, 1) = c + y * y * v;     R(1, 2) = y * z * v - x * s;
    R(2, 0) = z * x * v - y * s; R(2, 1) = z * y * v + x * s; R(2, 2) = c + z * z * v;
    return R;
  }
 private:
  T a_; Matrix<T, 3, 1> k_;
};
typedef AngleAxis<double> AngleAxisd; typedef AngleAxis<float> AngleAxisf;

// Transform<T,3,Isometry>: column-major 4x4, same bytes as Eigen's.
template <typename T, int Dim, int Mode>
class Transform {
  static_assert(Dim == 3, "only 3-D transforms");
 public:
  typedef Matrix<T, 3, 3> LinearMatrix;
  typedef Matrix<T, 3, 1> Vector;
  class LinearRef {
   public:
    explicit LinearRef(T *d) : d_(d) {}
    LinearRef &operator=(const LinearMatrix &R) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) d_[c * 4 + r] = R(r, c); return *this; }
    LinearRef &operator=(const LinearRef &o) { return *this = LinearMatrix(o); }
    operator LinearMatrix() const { LinearMatrix R; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = d_[c * 4 + r]; return R; }
    T &operator()(int r, int c) { return d_[c * 4 + r]; }
    T operator()(int r, int c) const { return d_[c * 4 + r]; }
    LinearMatrix operator*(const LinearMatrix &o) const { return LinearMatrix(*this) * o; }
    Vector operator*(const Vector &v) const { return LinearMatrix(*this) * v; }
    T norm() const { return LinearMatrix(*this).norm(); }
   private:
    T *d_;
  };
  class TranslationRef {
   public:
    explicit TranslationRef(T *d) : d_(d) {}
    TranslationRef &operator=(const Vector &v) { d_[0] = v(0); d_[1] = v(1); d_[2] = v(2); return *this; }
    TranslationRef &operator=(const TranslationRef &o) { return *this = Vector(o); }
    TranslationRef &operator*=(T s) { d_[0] *= s; d_[1] *= s; d_[2] *= s; return *this; }
    TranslationRef &operator+=(const Vector &v) { d_[0] += v(0); d_[1] += v(1); d_[2] += v(2); return *this; }
    operator Vector() const { return Vector(d_[0], d_[1], d_[2]); }
    Vector operator*(T s) const { return Vector(d_[0] * s, d_[1] * s, d_[2] * s); }
    void setZero() { d_[0] = d_[1] = d_[2] = T(0); }
    T &x() { return d_[0]; } T &y() { return d_[1]; } T &z() { return d_[2]; }
    T &operator()(int i) { return d_[i]; }
    T operator()(int i) const { return d_[i]; }
   private:
    T *d_;
  };
  Transform() { for (int i = 0; i < 16; ++i) m_[i] = T(0); m_[15] = T(1); }
  static Transform Identity() { Transform t; t.m_[0] = t.m_[5] = t.m_[10] = T(1); return t; }
  void setIdentity() { *this = Identity(); }
  LinearRef linear() { return LinearRef(m_); }
  LinearMatrix linear() const { return LinearMatrix(LinearRef(const_cast<T *>(m_))); }
  LinearMatrix rotation() const { return linear(); }
  TranslationRef translation() { return TranslationRef(m_ + 12); }
  Vector translation() const { return Vector(m_[12], m_[13], m_[14]); }
  T *data() { return m_; } const T *data() const { return m_; }
  Matrix<T, 4, 4> matrix() const { Matrix<T, 4, 4> M; for (int i = 0; i < 16; ++i) M.data()[i] = m_[i]; return M; }
  Transform inverse() const {  // Isometry: (R^T, -R^T t)
    Transform q;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) q.m_[c * 4 + r] = m_[r * 4 + c];
    for (int r = 0; r < 3; ++r) q.m_[12 + r] = -(q.m_[0 * 4 + r] * m_[12] + q.m_[1 * 4 + r] * m_[13] + q.m_[2 * 4 + r] * m_[14]);
    return q;
  }
  Transform operator*(const Transform &o) const {
    Transform q;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 4; ++c) { T s = 0; for (int k = 0; k < 3; ++k) s += m_[k * 4 + r] * o.m_[c * 4 + k]; q.m_[c * 4 + r] = s; }
      q.m_[12 + r] += m_[12 + r];
    }
    return q;
  }
  Vector operator*(const Vector &v) const {
    Vector q;
    for (int r = 0; r < 3; ++r) q(r) = m_[0 * 4 + r] * v(0) + m_[1 * 4 + r] * v(1) + m_[2 * 4 + r] * v(2) + m_[12 + r];
    return q;
  }
 private:
  T m_[16];
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<float, 3, Isometry> Isometry3f;

}  // namespace Eigen
#endif  // BA_B200_HAVE_EIGEN
#endif  // BA_B200_EIGEN_SHIM_H_
