// =============================================================================
// oracle/ba_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (never shipped, never timed
// as the product).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// CPU restatement (plain C++17, single thread, no Eigen) of the reference's
// analytic bundle-adjustment hot path:
//   full BA   : core/full_bundle_adjustment_solver.cpp:72-206, 381-500, 503-628,
//               630-1044, 1046-1082; the Gauss-Newton branch and SolveByGradientDescent of
//               core/full_bundle_adjustment_solver_refactor.cpp:944-982, 1075-1367 (same
//               linearisation, different step / acceptance)
//   pose-only : core/pose_only_bundle_adjustment_solver.cpp:8-399 (6-DoF),
//               401-900 (planar 3-DoF), 907-1278 (JtJ helpers), 1280-1316 (se3
//               exp), 1338-1583 (warp, Jacobians, gradient/Hessian)
//   Eigen LDLT: third-party (Eigen, un-pinned `find_package(Eigen3)`), restated
//               from its published algorithm (LDLT.h: diagonal-pivoted unblocked
//               LDL^T; solve() zeroes components whose |D| <= numeric_limits::min()).
//
// PARITY UNPINNED: the reference cannot be compiled in this container (Eigen,
// Ceres, OpenCV absent) and its tests hold no golden vectors for the solvers;
// the only known-answer data (test/test_projection_of_3d_point.cc:11-32) pins
// the projection convention and is checked in tests/.  The oracle is
// additionally cross-checked against finite differences and SciPy in tests/.
//
// Deviations from the reference that do not change arithmetic:
//   * parameters are keyed by insertion index, not by pointer / hash order
//     (reference numbers them in unordered_map order, which is unspecified);
//   * the dense N_opt x M_opt block arrays B_, Bt_, BCinv_, CinvBt_
//     (full...cpp:243-308) are stored sparsely (one block per observed pair);
//   * unqualified abs() is written std::fabs (see SURVEY 8c trap 1).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Eigen::LDLT restatement (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::
// unblocked + LDLT::_solve_impl).  Column-major n x n, lower triangle is used.
// ---------------------------------------------------------------------------
template <typename T>
struct Ldlt {
  int n = 0;
  std::vector<T> m;      // column-major, in-place factor (L strictly lower, D on diagonal)
  std::vector<int> tr;   // transpositions
  T &at(int r, int c) { return m[(size_t)c * n + r]; }

  void compute(int n_, const T *a_colmajor) {
    n = n_;
    m.assign(a_colmajor, a_colmajor + (size_t)n * n);
    tr.assign(n, 0);
    std::vector<T> temp(n);
    if (n <= 1) {
      if (n == 1) tr[0] = 0;
      return;
    }
    for (int k = 0; k < n; ++k) {
      // largest |diagonal| in the trailing corner (first maximum on ties)
      int big = k;
      T best = std::fabs(at(k, k));
      for (int i = k + 1; i < n; ++i) {
        T v = std::fabs(at(i, i));
        if (v > best) { best = v; big = i; }
      }
      tr[k] = big;
      if (k != big) {
        // symmetric row/column interchange on the lower triangle
        for (int c = 0; c < k; ++c) std::swap(at(k, c), at(big, c));
        for (int r = big + 1; r < n; ++r) std::swap(at(r, k), at(r, big));
        std::swap(at(k, k), at(big, big));
        for (int i = k + 1; i < big; ++i) std::swap(at(i, k), at(big, i));
      }
      const int rs = n - k - 1;
      if (k > 0) {
        // temp = D(0:k) .* A10^T ; A(k,k) -= A10 * temp ; A21 -= A20 * temp
        T acc = T(0);
        for (int c = 0; c < k; ++c) {
          temp[c] = at(c, c) * at(k, c);
          acc += at(k, c) * temp[c];
        }
        at(k, k) -= acc;
        if (rs > 0) {
          for (int c = 0; c < k; ++c) {
            const T tc = temp[c];
            if (tc == T(0)) continue;
            T *col = &m[(size_t)c * n];
            T *dst = &m[(size_t)k * n];
            for (int r = k + 1; r < n; ++r) dst[r] -= col[r] * tc;
          }
        }
      }
      const T akk = at(k, k);
      const bool pivot_is_valid = std::fabs(akk) > T(0);
      if (rs > 0 && pivot_is_valid) {
        T *dst = &m[(size_t)k * n];
        for (int r = k + 1; r < n; ++r) dst[r] /= akk;
      }
    }
  }

  // x <- A^{-1} b for nrhs right-hand sides stored column-major (n x nrhs)
  void solve(T *b, int nrhs) const {
    const T tol = std::numeric_limits<T>::min();
    for (int q = 0; q < nrhs; ++q) {
      T *v = b + (size_t)q * n;
      for (int k = 0; k < n; ++k)
        if (tr[k] != k) std::swap(v[k], v[tr[k]]);  // P b
      for (int c = 0; c < n; ++c) {                    // L^{-1}
        const T vc = v[c];
        if (vc == T(0)) continue;
        const T *col = &m[(size_t)c * n];
        for (int r = c + 1; r < n; ++r) v[r] -= col[r] * vc;
      }
      for (int i = 0; i < n; ++i) {                    // pseudo-inverse of D
        const T d = m[(size_t)i * n + i];
        if (std::fabs(d) > tol) v[i] /= d; else v[i] = T(0);
      }
      for (int r = n - 1; r >= 0; --r) {               // L^{-T}
        const T *col = &m[(size_t)r * n];
        T acc = v[r];
        for (int i = r + 1; i < n; ++i) acc -= col[i] * v[i];
        v[r] = acc;
      }
      for (int k = n - 1; k >= 0; --k)
        if (tr[k] != k) std::swap(v[k], v[tr[k]]);  // P^T
    }
  }
};

// ---------------------------------------------------------------------------
// small helpers (row-major 3x3 rotation + translation)
// ---------------------------------------------------------------------------
template <typename T>
struct PoseT {
  T R[9];
  T t[3];
};
template <typename T>
static PoseT<T> pose_identity() {
  PoseT<T> p;
  for (int i = 0; i < 9; ++i) p.R[i] = T(i % 4 == 0);
  p.t[0] = p.t[1] = p.t[2] = T(0);
  return p;
}
// Eigen::Transform<.,3,Isometry>::inverse(): (R^T, -R^T t)
template <typename T>
static PoseT<T> pose_inverse(const PoseT<T> &p) {
  PoseT<T> q;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) q.R[r * 3 + c] = p.R[c * 3 + r];
  for (int r = 0; r < 3; ++r)
    q.t[r] = -(q.R[r * 3 + 0] * p.t[0] + q.R[r * 3 + 1] * p.t[1] + q.R[r * 3 + 2] * p.t[2]);
  return q;
}
// a * b  (apply b first)
template <typename T>
static PoseT<T> pose_mul(const PoseT<T> &a, const PoseT<T> &b) {
  PoseT<T> q;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      q.R[r * 3 + c] = a.R[r * 3 + 0] * b.R[0 * 3 + c] + a.R[r * 3 + 1] * b.R[1 * 3 + c] +
                       a.R[r * 3 + 2] * b.R[2 * 3 + c];
    q.t[r] = a.R[r * 3 + 0] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
  return q;
}
template <typename T>
static void pose_apply(const PoseT<T> &p, const T *x, T *out) {
  for (int r = 0; r < 3; ++r)
    out[r] = p.R[r * 3 + 0] * x[0] + p.R[r * 3 + 1] * x[1] + p.R[r * 3 + 2] * x[2] + p.t[r];
}
// 4x4 column-major (Eigen::Transform::data()) -> PoseT
template <typename T>
static PoseT<T> pose_from_colmajor44(const T *d) {
  PoseT<T> p;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.R[r * 3 + c] = d[c * 4 + r];
    p.t[r] = d[12 + r];
  }
  return p;
}
template <typename T>
static void pose_to_colmajor44(const PoseT<T> &p, T *d) {
  for (int i = 0; i < 16; ++i) d[i] = T(0);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) d[c * 4 + r] = p.R[r * 3 + c];
    d[12 + r] = p.t[r];
  }
  d[15] = T(1);
}

// se3Exp: full...cpp:1046-1082 / pose_only...cpp:1280-1316.  xi = [v; w].
template <typename T>
static PoseT<T> se3_exp(const T *xi) {
  const T v[3] = {xi[0], xi[1], xi[2]};
  const T w[3] = {xi[3], xi[4], xi[5]};
  const T theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const T wx[9] = {T(0), -w[2], w[1], w[2], T(0), -w[0], -w[1], w[0], T(0)};
  T wx2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      wx2[r * 3 + c] = wx[r * 3 + 0] * wx[0 * 3 + c] + wx[r * 3 + 1] * wx[1 * 3 + c] +
                       wx[r * 3 + 2] * wx[2 * 3 + c];
  T R[9], V[9];
  if (theta < 1e-7) {
    for (int i = 0; i < 9; ++i) {
      const T I = T(i % 4 == 0);
      R[i] = I + wx[i] + T(0.5) * wx2[i];
      V[i] = I + T(0.5) * wx[i] + wx2[i] * T(0.33333333333333333333333333);
    }
  } else {
    const T s = std::sin(theta), c = std::cos(theta);
    const T a = s / theta;
    const T b = (T(1) - c) / (theta * theta);
    const T g = (theta - s) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) {
      const T I = T(i % 4 == 0);
      R[i] = I + a * wx[i] + b * wx2[i];
      V[i] = I + b * wx[i] + g * wx2[i];
    }
  }
  PoseT<T> p;
  for (int i = 0; i < 9; ++i) p.R[i] = R[i];
  for (int r = 0; r < 3; ++r) p.t[r] = V[r * 3 + 0] * v[0] + V[r * 3 + 1] * v[1] + V[r * 3 + 2] * v[2];
  return p;
}

// ===========================================================================
// Full bundle adjustment
// ===========================================================================
struct Camera {
  double fx, fy, cx, cy;
  PoseT<double> T_cj;  // pose_this_to_cam0, used as T_c<-body (full...cpp:746-747)
};
struct Observation {
  int cam, pose, point;
  double u, v;
};

struct IterInfo {  // mirrors OptimizationInfo (solver_option_and_summary.h:37-46)
  double cost, cost_change, average_reprojection_error, abs_gradient, abs_step, damping_term,
      iter_time;
  int iteration_status;
  int _pad;
};

struct FullOptions {  // mirrors Options (solver_option_and_summary.h:55-71), floats kept float
  int solver_type;
  float threshold_step_size, threshold_cost_change;
  float threshold_huber_loss, threshold_outlier_rejection;
  int max_num_iterations;
  float initial_lambda, decrease_ratio_lambda, increase_ratio_lambda;
  int b_accumulate;  // 0 = reference-exact (B assignment, last writer wins), 1 = corrected (B +=)
  int method;        // 0 = Solve (LM); FullBundleAdjustmentSolverRefactor: 1 = Gauss-Newton branch of Solve
                     // (full_bundle_adjustment_solver_refactor.cpp:976-982), 2 = SolveByGradientDescent (:1075-1367)
};

struct FullBA {
  double scaler = 0.01, inverse_scaler = 1.0 / 0.01;  // full...cpp:38-39
  std::unordered_map<int, Camera> cams;
  std::vector<PoseT<double>> T_jw;
  std::vector<char> pose_fixed;
  std::vector<double> X;  // 3 per point
  std::vector<char> point_fixed;
  std::vector<Observation> obs;

  // finalize
  bool finalized = false;
  int N = 0, M = 0;
  std::vector<int> pose_opt, point_opt, opt_pose, opt_point;
  std::vector<int> obs_pair;            // per observation: pair index or -1
  std::vector<int> point_pair_ptr;      // M+1
  std::vector<int> pair_pose;           // j_opt per pair, ascending within a point
  std::vector<int> pair_point;          // i_opt per pair
  // block storage
  std::vector<double> A, a, C, b, Cinv, Cinv_b, B, BCinv, BCinv_b, S, rhs, x, y;
  // last-iteration scalars
  double last_cost_prev = 0, last_cost_new = 0, last_model = 0, last_rho = 0, last_lambda = 0;
  double initial_cost = 0;
  int converged = 0;
  double t_linearize = 0, t_schur = 0, t_solve = 0, t_rest = 0;

  void add_camera(int id, double fx, double fy, double cx, double cy, const PoseT<double> &T) {
    // full...cpp:72-85 ; duplicate ids ignored by unordered_map::insert
    Camera c;
    c.fx = fx * scaler; c.fy = fy * scaler; c.cx = cx * scaler; c.cy = cy * scaler;
    c.T_cj = T;
    for (int k = 0; k < 3; ++k) c.T_cj.t[k] *= scaler;
    cams.insert({id, c});
  }
  int add_pose(const PoseT<double> &pose_cam_to_world) {
    // full...cpp:87-101
    PoseT<double> T = pose_inverse(pose_cam_to_world);
    for (int k = 0; k < 3; ++k) T.t[k] = T.t[k] * scaler;
    T_jw.push_back(T);
    pose_fixed.push_back(0);
    return (int)T_jw.size() - 1;
  }
  int add_point(const double *p) {
    // full...cpp:103-117
    for (int k = 0; k < 3; ++k) X.push_back(p[k] * scaler);
    point_fixed.push_back(0);
    return (int)point_fixed.size() - 1;
  }
  int add_observation(int cam, int pose, int point, double u, double v) {
    // full...cpp:155-180 ; invalid keys drop the observation
    if (cams.count(cam) == 0) return -1;
    if (pose < 0 || pose >= (int)T_jw.size()) return -2;
    if (point < 0 || point >= (int)point_fixed.size()) return -3;
    obs.push_back({cam, pose, point, u * scaler, v * scaler});
    return 0;
  }

  void finalize() {
    // full...cpp:182-206 (indices in insertion order) + connectivity (:669-700)
    if (finalized) return;
    pose_opt.assign(T_jw.size(), -1);
    point_opt.assign(point_fixed.size(), -1);
    opt_pose.clear(); opt_point.clear();
    for (size_t j = 0; j < T_jw.size(); ++j)
      if (!pose_fixed[j]) { pose_opt[j] = (int)opt_pose.size(); opt_pose.push_back((int)j); }
    for (size_t i = 0; i < point_fixed.size(); ++i)
      if (!point_fixed[i]) { point_opt[i] = (int)opt_point.size(); opt_point.push_back((int)i); }
    N = (int)opt_pose.size();
    M = (int)opt_point.size();
    // distinct (point, pose) pairs among free/free observations, sorted by (i_opt, j_opt)
    std::vector<uint64_t> keys;
    keys.reserve(obs.size());
    for (const auto &o : obs) {
      const int j = pose_opt[o.pose], i = point_opt[o.point];
      if (j >= 0 && i >= 0) keys.push_back(((uint64_t)i << 32) | (uint32_t)j);
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const size_t P = keys.size();
    pair_pose.resize(P); pair_point.resize(P);
    point_pair_ptr.assign(M + 1, 0);
    for (size_t p = 0; p < P; ++p) {
      pair_point[p] = (int)(keys[p] >> 32);
      pair_pose[p] = (int)(keys[p] & 0xffffffffu);
      point_pair_ptr[pair_point[p] + 1]++;
    }
    for (int i = 0; i < M; ++i) point_pair_ptr[i + 1] += point_pair_ptr[i];
    obs_pair.assign(obs.size(), -1);
    for (size_t k = 0; k < obs.size(); ++k) {
      const int j = pose_opt[obs[k].pose], i = point_opt[obs[k].point];
      if (j < 0 || i < 0) continue;
      const uint64_t key = ((uint64_t)i << 32) | (uint32_t)j;
      obs_pair[k] = (int)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
    }
    A.assign((size_t)N * 36, 0); a.assign((size_t)N * 6, 0);
    C.assign((size_t)M * 9, 0); b.assign((size_t)M * 3, 0);
    Cinv.assign((size_t)M * 9, 0); Cinv_b.assign((size_t)M * 3, 0);
    B.assign(P * 18, 0); BCinv.assign(P * 18, 0);
    BCinv_b.assign((size_t)N * 6, 0);
    S.assign((size_t)36 * N * N, 0); rhs.assign((size_t)6 * N, 0); x.assign((size_t)6 * N, 0);
    y.assign((size_t)M * 3, 0);
    finalized = true;
  }

  // projection residual shared by cost and linearisation (full...cpp:403-425 / 733-760)
  inline void project(const Observation &o, double *Xij, double *Xijc, double *rij,
                      const Camera **cam_out) const {
    const Camera &cam = cams.at(o.cam);
    const PoseT<double> &T = T_jw[o.pose];
    pose_apply(T, &X[(size_t)o.point * 3], Xij);
    pose_apply(cam.T_cj, Xij, Xijc);
    const double invz = 1.0 / Xijc[2];
    const double xinvz = Xijc[0] * invz, yinvz = Xijc[1] * invz;
    rij[0] = cam.fx * xinvz + cam.cx - o.u;
    rij[1] = cam.fy * yinvz + cam.cy - o.v;
    *cam_out = &cam;
  }

  double evaluate_current_cost() const {  // full...cpp:381-433
    double err = 0.0;
    for (const auto &o : obs) {
      double Xij[3], Xijc[3], r[2];
      const Camera *cam;
      project(o, Xij, Xijc, r, &cam);
      err += std::sqrt(r[0] * r[0] + r[1] * r[1]);
    }
    return err;
  }

  void linearize(float thres_huber, int b_accumulate) {  // full...cpp:711-831
    std::fill(A.begin(), A.end(), 0.0); std::fill(a.begin(), a.end(), 0.0);
    std::fill(C.begin(), C.end(), 0.0); std::fill(b.begin(), b.end(), 0.0);
    std::fill(B.begin(), B.end(), 0.0);
    for (size_t k = 0; k < obs.size(); ++k) {
      const Observation &o = obs[k];
      const Camera &cam = cams.at(o.cam);
      const int j_opt = pose_opt[o.pose], i_opt = point_opt[o.point];
      const PoseT<double> &T = T_jw[o.pose];
      double Xij[3], Xc[3];
      pose_apply(T, &X[(size_t)o.point * 3], Xij);
      pose_apply(cam.T_cj, Xij, Xc);
      const double invz = 1.0 / Xc[2];
      const double fxinvz = cam.fx * invz, fyinvz = cam.fy * invz;
      const double xinvz = Xc[0] * invz, yinvz = Xc[1] * invz;
      const double fx_xinvz2 = fxinvz * xinvz, fy_yinvz2 = fyinvz * yinvz;
      const double r0 = cam.fx * xinvz + cam.cx - o.u;
      const double r1 = cam.fy * yinvz + cam.cy - o.v;
      const double absrxry = std::fabs(r0) + std::fabs(r1);
      const double weight = (absrxry > thres_huber) ? (thres_huber / absrxry) : 1.0f;
      const double wr0 = weight * r0, wr1 = weight * r1;
      const double D00 = fxinvz, D02 = -fx_xinvz2, D11 = fyinvz, D12 = -fy_yinvz2;
      const double *Rc = cam.T_cj.R;
      double G[6];
      for (int c = 0; c < 3; ++c) {
        G[c] = D00 * Rc[0 * 3 + c] + D02 * Rc[2 * 3 + c];
        G[3 + c] = D11 * Rc[1 * 3 + c] + D12 * Rc[2 * 3 + c];
      }
      double Q[12];  // 2x6 row-major
      if (j_opt >= 0) {
        const double K[9] = {0.0, Xij[2], -Xij[1], -Xij[2], 0.0, Xij[0], Xij[1], -Xij[0], 0.0};
        for (int r = 0; r < 2; ++r) {
          for (int c = 0; c < 3; ++c) {
            Q[r * 6 + c] = G[r * 3 + c];
            Q[r * 6 + 3 + c] =
                G[r * 3 + 0] * K[0 * 3 + c] + G[r * 3 + 1] * K[1 * 3 + c] + G[r * 3 + 2] * K[2 * 3 + c];
          }
        }
        double *Aj = &A[(size_t)j_opt * 36];
        for (int r = 0; r < 6; ++r)
          for (int c = r; c < 6; ++c)
            Aj[r * 6 + c] += (weight * Q[r]) * Q[c] + (weight * Q[6 + r]) * Q[6 + c];
        double *aj = &a[(size_t)j_opt * 6];
        for (int r = 0; r < 6; ++r) aj[r] -= Q[r] * wr0 + Q[6 + r] * wr1;
      }
      if (i_opt >= 0) {
        double Rm[6];  // 2x3 = G * R_jw
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c)
            Rm[r * 3 + c] = G[r * 3 + 0] * T.R[0 * 3 + c] + G[r * 3 + 1] * T.R[1 * 3 + c] +
                            G[r * 3 + 2] * T.R[2 * 3 + c];
        double *Ci = &C[(size_t)i_opt * 9];
        for (int r = 0; r < 3; ++r)
          for (int c = r; c < 3; ++c)
            Ci[r * 3 + c] += weight * (Rm[r] * Rm[c] + Rm[3 + r] * Rm[3 + c]);
        double *bi = &b[(size_t)i_opt * 3];
        for (int r = 0; r < 3; ++r) bi[r] -= Rm[r] * wr0 + Rm[3 + r] * wr1;
        if (j_opt >= 0) {
          double *Bp = &B[(size_t)obs_pair[k] * 18];  // 6x3 row-major
          for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 3; ++c) {
              const double v = weight * (Q[r] * Rm[c] + Q[6 + r] * Rm[3 + c]);
              if (b_accumulate) Bp[r * 3 + c] += v; else Bp[r * 3 + c] = v;  // :826 assignment
            }
        }
      }
    }
  }

  void damp_and_invert(double lambda) {  // full...cpp:833-856
    const double lp1 = 1.0 + lambda;
    for (int j = 0; j < N; ++j) {
      double *Aj = &A[(size_t)j * 36];
      for (int r = 0; r < 6; ++r)
        for (int c = r + 1; c < 6; ++c) Aj[c * 6 + r] = Aj[r * 6 + c];
      for (int d = 0; d < 6; ++d) Aj[d * 6 + d] *= lp1;
    }
    Ldlt<double> ldlt;
    for (int i = 0; i < M; ++i) {
      double *Ci = &C[(size_t)i * 9];
      Ci[3] = Ci[1]; Ci[6] = Ci[2]; Ci[7] = Ci[5];
      Ci[0] *= lp1; Ci[4] *= lp1; Ci[8] *= lp1;
      ldlt.compute(3, Ci);  // symmetric: row-major == column-major
      double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      ldlt.solve(I, 3);     // columns of the inverse
      double *Ii = &Cinv[(size_t)i * 9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ii[r * 3 + c] = I[c * 3 + r];
      const double *bi = &b[(size_t)i * 3];
      for (int r = 0; r < 3; ++r)
        Cinv_b[(size_t)i * 3 + r] = Ii[r * 3 + 0] * bi[0] + Ii[r * 3 + 1] * bi[1] + Ii[r * 3 + 2] * bi[2];
    }
  }

  void schur() {  // full...cpp:858-888 + assembly :890-902 ; S stored column-major n x n
    const int n = 6 * N;
    std::fill(S.begin(), S.end(), 0.0);
    std::fill(BCinv_b.begin(), BCinv_b.end(), 0.0);
    // BCinvBt accumulated (blocks k >= j) straight into S's storage, negated at the end
    for (int i = 0; i < M; ++i) {
      const double *Ii = &Cinv[(size_t)i * 9];
      const double *bi = &b[(size_t)i * 3];
      for (int p = point_pair_ptr[i]; p < point_pair_ptr[i + 1]; ++p) {
        const double *Bp = &B[(size_t)p * 18];
        double *E = &BCinv[(size_t)p * 18];
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 3; ++c)
            E[r * 3 + c] = Bp[r * 3 + 0] * Ii[0 * 3 + c] + Bp[r * 3 + 1] * Ii[1 * 3 + c] +
                           Bp[r * 3 + 2] * Ii[2 * 3 + c];
        const int j = pair_pose[p];
        for (int r = 0; r < 6; ++r)
          BCinv_b[(size_t)j * 6 + r] += E[r * 3 + 0] * bi[0] + E[r * 3 + 1] * bi[1] + E[r * 3 + 2] * bi[2];
        for (int q = p; q < point_pair_ptr[i + 1]; ++q) {  // k >= j (pairs ascending in j)
          const int k = pair_pose[q];
          const double *Bq = &B[(size_t)q * 18];
          for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
              const double v = E[r * 3 + 0] * Bq[c * 3 + 0] + E[r * 3 + 1] * Bq[c * 3 + 1] +
                               E[r * 3 + 2] * Bq[c * 3 + 2];
              S[(size_t)(6 * k + c) * n + (6 * j + r)] += v;  // element (6j+r, 6k+c)
            }
        }
      }
    }
    // mirror (BCinvBt_[k][j] = BCinvBt_[j][k]^T, :874-876), S = A - BCinvBt (:878-885)
    for (int j = 0; j < N; ++j)
      for (int k = j; k < N; ++k)
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) {
            const size_t jk = (size_t)(6 * k + c) * n + (6 * j + r);  // (6j+r, 6k+c)
            const size_t kj = (size_t)(6 * j + r) * n + (6 * k + c);  // (6k+c, 6j+r)
            if (j == k && c < r) continue;  // diagonal block: keep upper, mirror below
            S[kj] = S[jk];
          }
    for (size_t e = 0; e < S.size(); ++e) S[e] = -S[e];
    for (int j = 0; j < N; ++j)
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c)
          S[(size_t)(6 * j + c) * n + (6 * j + r)] += A[(size_t)j * 36 + r * 6 + c];
    for (int j = 0; j < N; ++j)
      for (int r = 0; r < 6; ++r) rhs[(size_t)j * 6 + r] = a[(size_t)j * 6 + r] - BCinv_b[(size_t)j * 6 + r];
  }

  void solve_reduced() {  // full...cpp:890-917
    const int n = 6 * N;
    Ldlt<double> ldlt;
    ldlt.compute(n, S.data());
    x = rhs;
    if (n > 0) ldlt.solve(x.data(), 1);
    for (int i = 0; i < M; ++i) {
      double acc[3] = {0, 0, 0};
      for (int p = point_pair_ptr[i]; p < point_pair_ptr[i + 1]; ++p) {
        const double *E = &BCinv[(size_t)p * 18];  // CinvBt = BCinv^T
        const double *xj = &x[(size_t)pair_pose[p] * 6];
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 6; ++r) acc[c] += E[r * 3 + c] * xj[r];
      }
      for (int c = 0; c < 3; ++c) y[(size_t)i * 3 + c] = Cinv_b[(size_t)i * 3 + c] - acc[c];
    }
  }

  double model_change() const {  // full...cpp:435-455 (damped A, C)
    double e = 0.0;
    for (int j = 0; j < N; ++j) {
      const double *xj = &x[(size_t)j * 6], *aj = &a[(size_t)j * 6], *Aj = &A[(size_t)j * 36];
      for (int r = 0; r < 6; ++r) e += aj[r] * xj[r];
      double q = 0.0;
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int c = 0; c < 6; ++c) s += Aj[r * 6 + c] * xj[c];
        q += xj[r] * s;
      }
      e += q;
    }
    for (int i = 0; i < M; ++i) {
      const double *yi = &y[(size_t)i * 3], *bi = &b[(size_t)i * 3], *Ci = &C[(size_t)i * 9];
      for (int r = 0; r < 3; ++r) e += bi[r] * yi[r];
      double q = 0.0;
      for (int r = 0; r < 3; ++r) q += yi[r] * (Ci[r * 3] * yi[0] + Ci[r * 3 + 1] * yi[1] + Ci[r * 3 + 2] * yi[2]);
      e += q;
      double Bx[3] = {0, 0, 0};
      for (int p = point_pair_ptr[i]; p < point_pair_ptr[i + 1]; ++p) {
        const double *Bp = &B[(size_t)p * 18];
        const double *xj = &x[(size_t)pair_pose[p] * 6];
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 6; ++r) Bx[c] += Bp[r * 3 + c] * xj[r];
      }
      e += 2.0 * (yi[0] * Bx[0] + yi[1] * Bx[1] + yi[2] * Bx[2]);
    }
    return -e;
  }

  int solve(const FullOptions &opt, IterInfo *infos, int cap) {  // full...cpp:630-1044
    finalize();
    const int max_iteration = opt.max_num_iterations;
    const float THRES_HUBER = opt.threshold_huber_loss;
    const float THRES_DELTA_XI = opt.threshold_step_size;
    const float THRES_DELTA_ERROR = opt.threshold_cost_change;
    const double num_observations = (double)obs.size();
    bool is_converged = false;
    double previous_cost = evaluate_current_cost();
    initial_cost = previous_cost;
    double lambda = opt.initial_lambda;
    int n_done = 0;
    std::vector<PoseT<double>> reserved_poses(N);
    std::vector<double> reserved_points((size_t)M * 3);
    for (int iteration = 0; iteration < max_iteration; ++iteration) {
      linearize(THRES_HUBER, opt.b_accumulate);
      if (opt.method == 2) {
        // SolveByGradientDescent (refactor.cpp:1271-1283): the step is the gradient itself, every block clipped
        // to a norm of 0.001
        x.assign((size_t)N * 6, 0.0);
        y.assign((size_t)M * 3, 0.0);
        for (int j = 0; j < N; ++j) {
          double s2 = 0;
          for (int r = 0; r < 6; ++r) { x[(size_t)j * 6 + r] = a[(size_t)j * 6 + r]; s2 += x[(size_t)j * 6 + r] * x[(size_t)j * 6 + r]; }
          const double nrm = std::sqrt(s2);
          if (nrm > 0.001) for (int r = 0; r < 6; ++r) x[(size_t)j * 6 + r] = x[(size_t)j * 6 + r] * (0.001 / nrm);
        }
        for (int i = 0; i < M; ++i) {
          double s2 = 0;
          for (int r = 0; r < 3; ++r) { y[(size_t)i * 3 + r] = b[(size_t)i * 3 + r]; s2 += y[(size_t)i * 3 + r] * y[(size_t)i * 3 + r]; }
          const double nrm = std::sqrt(s2);
          if (nrm > 0.001) for (int r = 0; r < 3; ++r) y[(size_t)i * 3 + r] = y[(size_t)i * 3 + r] * (0.001 / nrm);
        }
      } else {
        damp_and_invert(lambda);
        schur();
        solve_reduced();
      }
      // reserve (:457-469) / update (:484-500)
      for (int j = 0; j < N; ++j) reserved_poses[j] = T_jw[opt_pose[j]];
      for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) reserved_points[(size_t)i * 3 + c] = X[(size_t)opt_point[i] * 3 + c];
      for (int j = 0; j < N; ++j) {
        const PoseT<double> d = se3_exp<double>(&x[(size_t)j * 6]);
        T_jw[opt_pose[j]] = pose_mul(d, T_jw[opt_pose[j]]);
      }
      for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) X[(size_t)opt_point[i] * 3 + c] += y[(size_t)i * 3 + c];

      const double current_cost = evaluate_current_cost();
      const double changed_error_by_model = (opt.method == 2) ? 1.0 : model_change();
      const double rho = (current_cost - previous_cost) * inverse_scaler / changed_error_by_model;
      last_cost_prev = previous_cost; last_cost_new = current_cost;
      last_model = changed_error_by_model; last_rho = rho; last_lambda = lambda;

      int iter_status;
      if (opt.method != 0) {
        iter_status = 0;  // refactor.cpp:976-982 / 1285-1289: the step is always kept, lambda untouched
      } else if (rho > 0.25) {
        iter_status = 0;  // UPDATE
      } else {
        for (int j = 0; j < N; ++j) T_jw[opt_pose[j]] = reserved_poses[j];
        for (int i = 0; i < M; ++i)
          for (int c = 0; c < 3; ++c) X[(size_t)opt_point[i] * 3 + c] = reserved_points[(size_t)i * 3 + c];
        iter_status = 2;  // SKIPPED
      }
      if (opt.method != 0) {
      } else if (rho > 0.5) {
        lambda = std::max(1e-10, static_cast<double>(lambda * opt.decrease_ratio_lambda));
        iter_status = 1;  // UPDATE_TRUST_MORE
      } else if (rho <= 0.25) {
        lambda = std::min(100.0, static_cast<double>(lambda * opt.increase_ratio_lambda));
      }
      const double average_error = current_cost / num_observations;
      const double cost_change = std::fabs(current_cost - previous_cost);
      // gradient descent starts both sums at 0.01 (refactor.cpp:1296-1297)
      double step_pose = (opt.method == 2) ? 0.01 : 0.0, step_point = (opt.method == 2) ? 0.01 : 0.0;
      for (int j = 0; j < N; ++j) {
        double s = 0;
        for (int r = 0; r < 6; ++r) s += x[(size_t)j * 6 + r] * x[(size_t)j * 6 + r];
        step_pose += std::sqrt(s);
      }
      for (int i = 0; i < M; ++i) {
        double s = 0;
        for (int r = 0; r < 3; ++r) s += y[(size_t)i * 3 + r] * y[(size_t)i * 3 + r];
        step_point += std::sqrt(s);
      }
      const double total_step_size = step_point + step_pose;
      const double average_total_step_size = total_step_size / static_cast<double>(N + M);
      if (average_total_step_size < THRES_DELTA_XI || cost_change < THRES_DELTA_ERROR) is_converged = true;
      if (iteration >= max_iteration - 1) is_converged = false;
      if (infos != nullptr && n_done < cap) {
        IterInfo &I = infos[n_done];
        I.cost = current_cost; I.cost_change = cost_change;
        I.average_reprojection_error = average_error;
        I.abs_step = average_total_step_size; I.abs_gradient = 0;
        I.damping_term = lambda; I.iter_time = 0; I.iteration_status = iter_status; I._pad = 0;
        if (iter_status == 2) {
          I.cost = previous_cost; I.cost_change = 0;
          I.average_reprojection_error = std::sqrt(previous_cost / num_observations);
        }
      }
      ++n_done;
      previous_cost = current_cost;
      if (is_converged) break;
    }
    converged = is_converged ? 1 : 0;
    return n_done;
  }
};

// ===========================================================================
// Pose-only solvers (float32)
// ===========================================================================
struct PoseOnlyOptions {
  float threshold_step_size, threshold_cost_change;
  float threshold_huber_loss, threshold_outlier_rejection;
  int max_num_iterations;
};
struct PoseOnlyResult {
  int n_iterations;   // loop trips executed (incl. the one that converged)
  int converged;
  int success;        // false only on NaN (pose_only...cpp:159-167)
  int n_summary;      // OptimizationInfo rows pushed (the converging trip pushes none)
  float final_error;
  float final_step;
};

// pose_only...cpp:1350-1384
static inline void jac_res_6dof(const float *Xl, const float *px, float fx, float fy, float cx,
                                float cy, float *Ju, float *Jv, float *res) {
  const float inverse_z = 1.0f / Xl[2];
  const float x_inverse_z = Xl[0] * inverse_z;
  const float y_inverse_z = Xl[1] * inverse_z;
  const float fx_x_inverse_z = fx * x_inverse_z;
  const float fy_y_inverse_z = fy * y_inverse_z;
  res[0] = (fx_x_inverse_z + cx) - px[0];
  res[1] = (fy_y_inverse_z + cy) - px[1];
  Ju[0] = fx * inverse_z; Ju[1] = 0.0f; Ju[2] = -fx_x_inverse_z * inverse_z;
  Ju[3] = -fx_x_inverse_z * y_inverse_z; Ju[4] = fx * (1.0f + x_inverse_z * x_inverse_z);
  Ju[5] = -fx * y_inverse_z;
  Jv[0] = 0.0f; Jv[1] = fy * inverse_z; Jv[2] = -fy_y_inverse_z * inverse_z;
  Jv[3] = -fy * (1.0f + y_inverse_z * y_inverse_z); Jv[4] = fy_y_inverse_z * x_inverse_z;
  Jv[5] = fy * x_inverse_z;
}
// pose_only...cpp:1454-1515
static inline void jac_res_3dof(const float *Xl, const float *Xw, const float *px, float fx, float fy,
                                float cx, float cy, const float *Rcb /*row-major 3x3*/, float cos_psi,
                                float sin_psi, float *Ju, float *Jv, float *res) {
  const float r11 = Rcb[0], r12 = Rcb[1], r21 = Rcb[3], r22 = Rcb[4], r31 = Rcb[6], r32 = Rcb[7];
  const float inverse_z = 1.0f / Xl[2];
  const float x_inverse_z = Xl[0] * inverse_z;
  const float y_inverse_z = Xl[1] * inverse_z;
  const float fx_x_inverse_z = fx * x_inverse_z;
  const float fy_y_inverse_z = fy * y_inverse_z;
  res[0] = (fx_x_inverse_z + cx) - px[0];
  res[1] = (fy_y_inverse_z + cy) - px[1];
  const float alpha_1 = fx * inverse_z, alpha_2 = -fx_x_inverse_z * inverse_z;
  const float beta_1 = fy * inverse_z, beta_2 = -fy_y_inverse_z * inverse_z;
  const float xb = Xw[0], yb = Xw[1];
  const float Aa = -sin_psi * xb - cos_psi * yb;
  const float Bb = cos_psi * xb - sin_psi * yb;
  Ju[0] = alpha_1 * r11 + alpha_2 * r31;
  Ju[1] = alpha_1 * r12 + alpha_2 * r32;
  Ju[2] = Ju[0] * Aa + Ju[1] * Bb;
  Jv[0] = beta_1 * r21 + beta_2 * r31;
  Jv[1] = beta_1 * r22 + beta_2 * r32;
  Jv[2] = Jv[0] * Aa + Jv[1] * Bb;
}
// pose_only...cpp:1386-1452 / 1516-1583 (D = 6 or 3).  H upper triangle, g, error quirk.
template <int D>
static inline void grad_hess(const float *Ju, const float *Jv, const float *res, float thres_huber,
                             float *H /*DxD row-major, upper*/, float *mJtWr, float &err_curr,
                             float &error_nonweighted) {
  const float abs_residual_sum = std::fabs(res[0]) + std::fabs(res[1]);
  error_nonweighted = abs_residual_sum;
  const float ru = res[0], rv = res[1];
  float g[D];
  float error = 0.0f;
  if (abs_residual_sum >= thres_huber) {
    const float weight = thres_huber / abs_residual_sum;
    const float wru = weight * ru, wrv = weight * rv;
    for (int r = 0; r < D; ++r) {
      const float wJu = weight * Ju[r], wJv = weight * Jv[r];
      for (int c = r; c < D; ++c) H[r * D + c] += (wJu * Ju[c] + wJv * Jv[c]);
    }
    for (int r = 0; r < D; ++r) g[r] = wru * Ju[r] + wrv * Jv[r];
    error += wru * ru;  // only the u term (:1432 / :1563)
  } else {
    for (int r = 0; r < D; ++r)
      for (int c = r; c < D; ++c) H[r * D + c] += (Ju[r] * Ju[c] + Jv[r] * Jv[c]);
    for (int r = 0; r < D; ++r) g[r] = ru * Ju[r] + rv * Jv[r];
    error += rv * rv;   // only the v term (:1450 / :1581)
  }
  for (int r = 0; r < D; ++r) mJtWr[r] -= g[r];
  err_curr += error;
}

template <int D>
static void solve_small(float *H, const float *g, float lambda, float *delta) {
  for (int r = 0; r < D; ++r)
    for (int c = r + 1; c < D; ++c) H[c * D + r] = H[r * D + c];
  for (int i = 0; i < D; ++i) H[i * D + i] *= (1.0f + lambda);
  Ldlt<float> ldlt;
  ldlt.compute(D, H);
  for (int i = 0; i < D; ++i) delta[i] = g[i];
  ldlt.solve(delta, 1);
}

// kind: 0 mono 6dof, 1 stereo 6dof.  pose_io: reference_to_current (R row-major 9 | t 3).
static PoseOnlyResult poseonly_6dof(int stereo, int n_pts, const float *Xw, const float *pxl,
                                    const float *pxr, const float *intr_l, const float *intr_r,
                                    const PoseT<float> *left_to_right, PoseT<float> *pose_io,
                                    unsigned char *mask_l, unsigned char *mask_r,
                                    const PoseOnlyOptions &opt, float *hist_cost, float *hist_step,
                                    float *debug_poses) {
  PoseOnlyResult res{};
  const int MAX_ITERATION = opt.max_num_iterations;
  const float THRES_HUBER = opt.threshold_huber_loss;
  const float THRES_DELTA_XI = opt.threshold_step_size;
  const float THRES_DELTA_ERROR = opt.threshold_cost_change;
  const float THRES_REPROJ_ERROR = opt.threshold_outlier_rejection;
  const float inverse_n_pts = 1.0f / static_cast<float>(n_pts);
  for (int i = 0; i < n_pts; ++i) { mask_l[i] = 1; if (stereo) mask_r[i] = 1; }
  PoseT<float> T_rl = pose_identity<float>();
  if (stereo) T_rl = pose_inverse(*left_to_right);
  PoseT<float> T_cw = pose_inverse(*pose_io);
  bool is_converged = true;
  float err_prev = 1e10f;
  const float lambda = 1e-5f;
  float last_err = 0, last_step = 0;
  for (int iter = 0; iter < MAX_ITERATION; ++iter) {
    float H[36] = {0}, g[6] = {0};
    float err_curr = 0.0f;
    size_t count_left = 0, count_right = 0;
    for (int i = 0; i < n_pts; ++i) {
      float Xl[3], Xr[3], Ju[6], Jv[6], r2[2], enw;
      pose_apply(T_cw, &Xw[3 * i], Xl);
      jac_res_6dof(Xl, &pxl[2 * i], intr_l[0], intr_l[1], intr_l[2], intr_l[3], Ju, Jv, r2);
      grad_hess<6>(Ju, Jv, r2, THRES_HUBER, H, g, err_curr, enw);
      ++count_left;
      if (enw >= THRES_REPROJ_ERROR) mask_l[i] = 0;
      if (!stereo) continue;
      if (pxr[2 * i] < 0 || pxr[2 * i + 1] < 0) continue;
      ++count_right;
      pose_apply(T_rl, Xl, Xr);
      jac_res_6dof(Xr, &pxr[2 * i], intr_r[0], intr_r[1], intr_r[2], intr_r[3], Ju, Jv, r2);
      grad_hess<6>(Ju, Jv, r2, THRES_HUBER, H, g, err_curr, enw);
      if (enw >= THRES_REPROJ_ERROR) mask_r[i] = 0;
    }
    float delta[6];
    solve_small<6>(H, g, lambda, delta);
    const PoseT<float> dT = se3_exp<float>(delta);
    T_cw = pose_mul(dT, T_cw);
    if (debug_poses) {
      const PoseT<float> inv = pose_inverse(T_cw);
      std::memcpy(debug_poses + 12 * iter, inv.R, 9 * sizeof(float));
      std::memcpy(debug_poses + 12 * iter + 9, inv.t, 3 * sizeof(float));
    }
    if (stereo) err_curr /= (count_left + count_right) * 0.5f;
    else err_curr *= (inverse_n_pts * 0.5f);
    const float delta_error = std::fabs(err_curr - err_prev);
    float nrm = 0;
    for (int k = 0; k < 6; ++k) nrm += delta[k] * delta[k];
    nrm = std::sqrt(nrm);
    res.n_iterations = iter + 1;
    last_err = err_curr; last_step = nrm;
    if (nrm < THRES_DELTA_XI || delta_error < THRES_DELTA_ERROR) { is_converged = true; break; }
    if (iter == MAX_ITERATION - 1) is_converged = false;
    if (hist_cost) hist_cost[res.n_summary] = err_curr;
    if (hist_step) hist_step[res.n_summary] = nrm;
    res.n_summary++;
    err_prev = err_curr;
  }
  res.converged = is_converged;
  res.final_error = last_err; res.final_step = last_step;
  float nn = 0;
  for (int k = 0; k < 9; ++k) nn += T_cw.R[k] * T_cw.R[k];
  if (!std::isnan(std::sqrt(nn))) { *pose_io = pose_inverse(T_cw); res.success = 1; }
  else res.success = 0;
  return res;
}

// planar 3-DoF (pose_only...cpp:401-900).  pose_io = pose_world_to_current.
static PoseOnlyResult poseonly_3dof(int stereo, int n_pts, const float *Xw, const float *pxl,
                                    const float *pxr, const float *intr_l, const float *intr_r,
                                    const PoseT<float> &base_to_camera,
                                    const PoseT<float> *left_to_right,
                                    const PoseT<float> &world_to_last, PoseT<float> *pose_io,
                                    unsigned char *mask_l, unsigned char *mask_r,
                                    const PoseOnlyOptions &opt, float *hist_cost, float *hist_step,
                                    float *debug_poses) {
  PoseOnlyResult res{};
  const int MAX_ITERATION = opt.max_num_iterations;
  const float THRES_HUBER = opt.threshold_huber_loss;
  const float THRES_DELTA_XI = opt.threshold_step_size;
  const float THRES_DELTA_ERROR = opt.threshold_cost_change;
  const float THRES_REPROJ_ERROR = opt.threshold_outlier_rejection;
  const float inverse_n_pts = 1.0f / static_cast<float>(n_pts);
  for (int i = 0; i < n_pts; ++i) { mask_l[i] = 1; if (stereo) mask_r[i] = 1; }
  PoseT<float> T_rl = pose_identity<float>();
  if (stereo) T_rl = pose_inverse(*left_to_right);
  const PoseT<float> T_cb = pose_inverse(base_to_camera);  // pose_camera_to_base
  float R_cb_left[9], R_cb_right[9];
  for (int k = 0; k < 9; ++k) R_cb_left[k] = T_cb.R[k];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R_cb_right[r * 3 + c] = T_rl.R[r * 3 + 0] * T_cb.R[0 * 3 + c] + T_rl.R[r * 3 + 1] * T_cb.R[1 * 3 + c] +
                              T_rl.R[r * 3 + 2] * T_cb.R[2 * 3 + c];
  const PoseT<float> c2c1 = pose_mul(pose_inverse(*pose_io), world_to_last);
  const PoseT<float> b2b1 = pose_mul(pose_mul(base_to_camera, c2c1), T_cb);
  float prm[3] = {b2b1.t[0], b2b1.t[1], std::atan2(b2b1.R[3], b2b1.R[0])};
  PoseT<float> T_b2b1 = pose_identity<float>();
  PoseT<float> T_wc_opt = *pose_io;
  bool is_converged = true;
  float err_prev = 1e10f;
  const float lambda = 1e-5f;
  float last_err = 0, last_step = 0;
  for (int iter = 0; iter < MAX_ITERATION; ++iter) {
    float H[9] = {0}, g[3] = {0};
    const float cos_psi = std::cos(prm[2]), sin_psi = std::sin(prm[2]);
    T_b2b1 = pose_identity<float>();
    T_b2b1.R[0] = cos_psi; T_b2b1.R[1] = -sin_psi; T_b2b1.R[3] = sin_psi; T_b2b1.R[4] = cos_psi;
    T_b2b1.t[0] = prm[0]; T_b2b1.t[1] = prm[1]; T_b2b1.t[2] = 0;
    const PoseT<float> T_left = pose_mul(T_cb, T_b2b1);
    const PoseT<float> T_right = pose_mul(T_rl, T_left);
    float err_curr = 0.0f;
    size_t count_left = 0, count_right = 0;
    for (int i = 0; i < n_pts; ++i) {
      float Xl[3], Xr[3], Ju[3], Jv[3], r2[2], enw;
      pose_apply(T_left, &Xw[3 * i], Xl);
      jac_res_3dof(Xl, &Xw[3 * i], &pxl[2 * i], intr_l[0], intr_l[1], intr_l[2], intr_l[3], R_cb_left,
                   cos_psi, sin_psi, Ju, Jv, r2);
      grad_hess<3>(Ju, Jv, r2, THRES_HUBER, H, g, err_curr, enw);
      ++count_left;
      if (enw >= THRES_REPROJ_ERROR) mask_l[i] = 0;
      if (!stereo) continue;
      if (pxr[2 * i] < 0 || pxr[2 * i + 1] < 0) continue;
      ++count_right;
      pose_apply(T_right, &Xw[3 * i], Xr);
      jac_res_3dof(Xr, &Xw[3 * i], &pxr[2 * i], intr_r[0], intr_r[1], intr_r[2], intr_r[3], R_cb_right,
                   cos_psi, sin_psi, Ju, Jv, r2);
      grad_hess<3>(Ju, Jv, r2, THRES_HUBER, H, g, err_curr, enw);
      if (enw >= THRES_REPROJ_ERROR) mask_r[i] = 0;
    }
    float delta[3];
    solve_small<3>(H, g, lambda, delta);
    PoseT<float> dT = pose_identity<float>();
    dT.R[0] = std::cos(delta[2]); dT.R[1] = -std::sin(delta[2]);
    dT.R[3] = std::sin(delta[2]); dT.R[4] = std::cos(delta[2]);
    dT.t[0] = delta[0]; dT.t[1] = delta[1]; dT.t[2] = 0;
    T_b2b1 = pose_mul(dT, T_b2b1);
    prm[0] = T_b2b1.t[0]; prm[1] = T_b2b1.t[1]; prm[2] += delta[2];
    T_wc_opt = pose_mul(pose_inverse(T_b2b1), base_to_camera);
    if (debug_poses) {
      std::memcpy(debug_poses + 12 * iter, T_wc_opt.R, 9 * sizeof(float));
      std::memcpy(debug_poses + 12 * iter + 9, T_wc_opt.t, 3 * sizeof(float));
    }
    if (stereo) err_curr /= (count_left + count_right) * 0.5f;
    else err_curr *= (inverse_n_pts * 0.5f);
    const float delta_error = std::fabs(err_curr - err_prev);
    const float nrm = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    res.n_iterations = iter + 1;
    last_err = err_curr; last_step = nrm;
    if (nrm < THRES_DELTA_XI || delta_error < THRES_DELTA_ERROR) { is_converged = true; break; }
    if (iter == MAX_ITERATION - 1) is_converged = false;
    if (hist_cost) hist_cost[res.n_summary] = err_curr;
    if (hist_step) hist_step[res.n_summary] = nrm;
    res.n_summary++;
    err_prev = err_curr;
  }
  res.converged = is_converged;
  res.final_error = last_err; res.final_step = last_step;
  float nn = 0;
  for (int k = 0; k < 9; ++k) nn += T_b2b1.R[k] * T_b2b1.R[k];
  if (!std::isnan(std::sqrt(nn))) { *pose_io = T_wc_opt; res.success = 1; }
  else res.success = 0;
  return res;
}

static PoseT<float> pose12f(const float *p) {
  PoseT<float> q;
  std::memcpy(q.R, p, 9 * sizeof(float));
  std::memcpy(q.t, p + 9, 3 * sizeof(float));
  return q;
}

}  // namespace

// ===========================================================================
// C interface (ctypes)
// ===========================================================================
extern "C" {

void *orc_full_create() { return new FullBA(); }
void orc_full_destroy(void *h) { delete static_cast<FullBA *>(h); }
void orc_full_add_camera(void *h, int id, double fx, double fy, double cx, double cy,
                         const double *T44_colmajor) {
  static_cast<FullBA *>(h)->add_camera(id, fx, fy, cx, cy, pose_from_colmajor44(T44_colmajor));
}
int orc_full_add_pose(void *h, const double *T44_colmajor) {
  return static_cast<FullBA *>(h)->add_pose(pose_from_colmajor44(T44_colmajor));
}
int orc_full_add_point(void *h, const double *X) { return static_cast<FullBA *>(h)->add_point(X); }
void orc_full_add_poses(void *h, int n, const double *T44s) {
  for (int k = 0; k < n; ++k) static_cast<FullBA *>(h)->add_pose(pose_from_colmajor44(T44s + 16 * k));
}
void orc_full_add_points(void *h, int n, const double *Xs) {
  for (int k = 0; k < n; ++k) static_cast<FullBA *>(h)->add_point(Xs + 3 * k);
}
void orc_full_make_pose_fixed(void *h, int id) { static_cast<FullBA *>(h)->pose_fixed[id] = 1; }
void orc_full_make_point_fixed(void *h, int id) { static_cast<FullBA *>(h)->point_fixed[id] = 1; }
int orc_full_add_observation(void *h, int cam, int pose, int point, double u, double v) {
  return static_cast<FullBA *>(h)->add_observation(cam, pose, point, u, v);
}
long long orc_full_add_observations(void *h, long long n, const int *cam, const int *pose,
                                    const int *point, const double *uv) {
  long long ok = 0;
  for (long long k = 0; k < n; ++k)
    ok += (static_cast<FullBA *>(h)->add_observation(cam[k], pose[k], point[k], uv[2 * k], uv[2 * k + 1]) == 0);
  return ok;
}
int orc_full_solve(void *h, const FullOptions *opt, IterInfo *infos, int cap) {
  return static_cast<FullBA *>(h)->solve(*opt, infos, cap);
}
int orc_full_converged(void *h) { return static_cast<FullBA *>(h)->converged; }
double orc_full_initial_cost(void *h) { return static_cast<FullBA *>(h)->initial_cost; }
// user-facing results (full...cpp:1011-1022): pose = (T_jw, t*100)^-1, point = X*100
void orc_full_get_pose(void *h, int id, double *T44_colmajor) {
  FullBA *s = static_cast<FullBA *>(h);
  PoseT<double> T = s->T_jw[id];
  for (int k = 0; k < 3; ++k) T.t[k] *= s->inverse_scaler;
  pose_to_colmajor44(pose_inverse(T), T44_colmajor);
}
void orc_full_get_point(void *h, int id, double *X) {
  FullBA *s = static_cast<FullBA *>(h);
  for (int k = 0; k < 3; ++k) X[k] = s->X[(size_t)id * 3 + k] * s->inverse_scaler;
}
// internal (scaled) state, R row-major 9 | t 3 per pose
void orc_full_get_internal(void *h, double *T_jw12, double *X3) {
  FullBA *s = static_cast<FullBA *>(h);
  for (size_t j = 0; j < s->T_jw.size(); ++j) {
    std::memcpy(T_jw12 + 12 * j, s->T_jw[j].R, 9 * sizeof(double));
    std::memcpy(T_jw12 + 12 * j + 9, s->T_jw[j].t, 3 * sizeof(double));
  }
  std::memcpy(X3, s->X.data(), s->X.size() * sizeof(double));
}
void orc_full_sizes(void *h, long long *out /*N, M, P, n_obs, N_total, M_total*/) {
  FullBA *s = static_cast<FullBA *>(h);
  s->finalize();
  out[0] = s->N; out[1] = s->M; out[2] = (long long)s->pair_pose.size();
  out[3] = (long long)s->obs.size(); out[4] = (long long)s->T_jw.size();
  out[5] = (long long)s->point_fixed.size();
}
// Dump of the block storage after the last executed iteration.
// which: 0 A(N*36) 1 a(N*6) 2 C(M*9) 3 b(M*3) 4 Cinv(M*9) 5 B(P*18) 6 S(n*n colmajor) 7 rhs 8 x 9 y
//        10 scalars {cost_prev,cost_new,model,rho,lambda}
long long orc_full_dump(void *h, int which, double *buf) {
  FullBA *s = static_cast<FullBA *>(h);
  const std::vector<double> *v = nullptr;
  switch (which) {
    case 0: v = &s->A; break; case 1: v = &s->a; break; case 2: v = &s->C; break;
    case 3: v = &s->b; break; case 4: v = &s->Cinv; break; case 5: v = &s->B; break;
    case 6: v = &s->S; break; case 7: v = &s->rhs; break; case 8: v = &s->x; break;
    case 9: v = &s->y; break;
    case 10:
      if (buf) { buf[0] = s->last_cost_prev; buf[1] = s->last_cost_new; buf[2] = s->last_model;
                 buf[3] = s->last_rho; buf[4] = s->last_lambda; }
      return 5;
    default: return -1;
  }
  if (buf) std::memcpy(buf, v->data(), v->size() * sizeof(double));
  return (long long)v->size();
}
void orc_full_pairs(void *h, int *pair_pose_id, int *pair_point_id) {  // original ids
  FullBA *s = static_cast<FullBA *>(h);
  for (size_t p = 0; p < s->pair_pose.size(); ++p) {
    pair_pose_id[p] = s->opt_pose[s->pair_pose[p]];
    pair_point_id[p] = s->opt_point[s->pair_point[p]];
  }
}
void orc_full_opt_ids(void *h, int *opt_pose_ids, int *opt_point_ids) {
  FullBA *s = static_cast<FullBA *>(h);
  std::copy(s->opt_pose.begin(), s->opt_pose.end(), opt_pose_ids);
  std::copy(s->opt_point.begin(), s->opt_point.end(), opt_point_ids);
}
// one linearisation + Schur build at the current parameters (for phase timing / shard tests):
// returns cost; fills internal blocks.  obs range [o0, o1) restricts nothing (whole problem).
double orc_full_build_only(void *h, float thres_huber, double lambda, int b_accumulate, int do_solve) {
  FullBA *s = static_cast<FullBA *>(h);
  s->finalize();
  s->linearize(thres_huber, b_accumulate);
  s->damp_and_invert(lambda);
  s->schur();
  if (do_solve) s->solve_reduced();
  return 0.0;
}
double orc_full_cost(void *h) { return static_cast<FullBA *>(h)->evaluate_current_cost(); }

// generic LDLT solve (double), column-major, for unit tests of the restatement
void orc_ldlt_solve_f64(int n, const double *A_colmajor, double *b, int nrhs) {
  Ldlt<double> l; l.compute(n, A_colmajor); l.solve(b, nrhs);
}
void orc_ldlt_solve_f32(int n, const float *A_colmajor, float *b, int nrhs) {
  Ldlt<float> l; l.compute(n, A_colmajor); l.solve(b, nrhs);
}
void orc_se3_exp_f64(const double *xi, double *Rt12) {
  PoseT<double> p = se3_exp<double>(xi);
  std::memcpy(Rt12, p.R, 9 * sizeof(double)); std::memcpy(Rt12 + 9, p.t, 3 * sizeof(double));
}

// Pose-only.  kind: 0 mono-6dof, 1 stereo-6dof, 2 mono-planar3dof, 3 stereo-planar3dof.
// poses are 12 floats (R row-major | t).  aux poses: left_to_right, base_to_camera, world_to_last.
// hist_* / debug_poses may be null; sized max_num_iterations (x12).
void orc_poseonly_solve(int kind, int n_pts, const float *Xw, const float *pxl, const float *pxr,
                        const float *intr_l, const float *intr_r, const float *left_to_right,
                        const float *base_to_camera, const float *world_to_last, float *pose_io,
                        unsigned char *mask_l, unsigned char *mask_r, const PoseOnlyOptions *opt,
                        PoseOnlyResult *result, float *hist_cost, float *hist_step,
                        float *debug_poses) {
  PoseT<float> pose = pose12f(pose_io);
  PoseT<float> l2r = left_to_right ? pose12f(left_to_right) : pose_identity<float>();
  PoseOnlyResult r;
  if (kind < 2) {
    r = poseonly_6dof(kind == 1, n_pts, Xw, pxl, pxr, intr_l, intr_r, &l2r, &pose, mask_l, mask_r, *opt,
                      hist_cost, hist_step, debug_poses);
  } else {
    r = poseonly_3dof(kind == 3, n_pts, Xw, pxl, pxr, intr_l, intr_r, pose12f(base_to_camera), &l2r,
                      pose12f(world_to_last), &pose, mask_l, mask_r, *opt, hist_cost, hist_step,
                      debug_poses);
  }
  std::memcpy(pose_io, pose.R, 9 * sizeof(float));
  std::memcpy(pose_io + 9, pose.t, 3 * sizeof(float));
  *result = r;
}
// batched convenience for CPU-baseline timing (frames independent, sequential)
void orc_poseonly_solve_batched(int kind, int n_frames, const int *offsets, const float *Xw,
                                const float *pxl, const float *pxr, const float *intr_l,
                                const float *intr_r, const float *left_to_right,
                                const float *base_to_camera, const float *world_to_last,
                                float *poses_io, unsigned char *mask_l, unsigned char *mask_r,
                                const PoseOnlyOptions *opt, PoseOnlyResult *results) {
  for (int f = 0; f < n_frames; ++f) {
    const int o = offsets[f], n = offsets[f + 1] - offsets[f];
    orc_poseonly_solve(kind, n, Xw + 3 * (size_t)o, pxl + 2 * (size_t)o, pxr ? pxr + 2 * (size_t)o : nullptr,
                       intr_l, intr_r, left_to_right, base_to_camera,
                       world_to_last ? world_to_last + 12 * (size_t)f : nullptr, poses_io + 12 * (size_t)f,
                       mask_l + o, mask_r ? mask_r + o : nullptr, opt, results + f, nullptr, nullptr,
                       nullptr);
  }
}

}  // extern "C"

// ===========================================================================
// utility/geometry_library.cpp restated (the checker of ba_geometry_batched and of the drop-in header
// ba_b200/utility/geometry_library.h).  Written the way the reference writes it: explicit 3 x 3 matrices, the
// skew matrix and its square formed and multiplied out.  Rotation matrices row-major, transforms R | t.
// ===========================================================================
namespace geom_ref {
template <typename T> struct M3 { T m[3][3]; };
template <typename T> M3<T> eye3() { M3<T> I{}; for (int i = 0; i < 3; ++i) I.m[i][i] = T(1); return I; }
template <typename T> M3<T> skew(T a, T b, T c) {       // :6-12  [w]x
  M3<T> S{};
  S.m[0][1] = -c; S.m[0][2] = b; S.m[1][0] = c; S.m[1][2] = -a; S.m[2][0] = -b; S.m[2][1] = a;
  return S;
}
template <typename T> M3<T> mul(const M3<T> &A, const M3<T> &B) {
  M3<T> C{};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) C.m[i][j] += A.m[i][k] * B.m[k][j];
  return C;
}
template <typename T> M3<T> lin(T a, const M3<T> &A, T b, const M3<T> &B, T c, const M3<T> &C) {
  M3<T> R{};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = a * A.m[i][j] + b * B.m[i][j] + c * C.m[i][j];
  return R;
}
template <typename T> void put(const M3<T> &A, T *o) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[3 * i + j] = A.m[i][j]; }
template <typename T> M3<T> get(const T *o) { M3<T> A; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A.m[i][j] = o[3 * i + j]; return A; }

template <typename T> void se3Exp(const T *xi, T *R9, T *t3) {   // :370-427
  const T v[3] = {xi[0], xi[1], xi[2]}, w[3] = {xi[3], xi[4], xi[5]};
  const T theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const M3<T> wx = skew(w[0], w[1], w[2]), wxwx = mul(wx, wx), I = eye3<T>();
  M3<T> R, V;
  if (theta < T(1e-9)) {
    R = lin(T(1), I, T(1), wx, T(0.5), wxwx);
    V = lin(T(1), I, T(0.5), wx, T(0.33333333333333333333333333), wxwx);
  } else {
    const T invtheta2 = T(1) / (theta * theta);
    R = lin(T(1), I, std::sin(theta) / theta, wx, (T(1) - std::cos(theta)) * invtheta2, wxwx);
    V = lin(T(1), I, (T(1) - std::cos(theta)) * invtheta2, wx, (theta - std::sin(theta)) / (theta * theta * theta), wxwx);
  }
  put(R, R9);
  for (int i = 0; i < 3; ++i) t3[i] = V.m[i][0] * v[0] + V.m[i][1] * v[1] + V.m[i][2] * v[2];
}
template <typename T> void so3Exp(const T *w, T *R9) {            // :590-611
  const T theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const M3<T> wx = skew(w[0], w[1], w[2]), wxwx = mul(wx, wx), I = eye3<T>();
  M3<T> R;
  if (theta < T(1e-9)) R = lin(T(1), I, T(1), wx, T(0.5), wxwx);
  else R = lin(T(1), I, std::sin(theta) / theta, wx, (T(1) - std::cos(theta)) * (T(1) / (theta * theta)), wxwx);
  put(R, R9);
}
template <typename T> bool so3LogCore(const M3<T> &R, T *w, T &theta) {   // :659-679 ; false: the small-angle branch
  const T inCos = (R.m[0][0] + R.m[1][1] + R.m[2][2] - T(1)) * T(0.5);
  theta = T(0);
  if (inCos >= T(0.999999999)) { w[0] = w[1] = w[2] = T(0); return false; }
  theta = std::acos(inCos);
  const T k = theta / (T(2) * std::sin(theta));
  M3<T> lnR;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lnR.m[i][j] = k * (R.m[i][j] - R.m[j][i]);
  w[0] = -lnR.m[1][2]; w[1] = lnR.m[0][2]; w[2] = -lnR.m[0][1];
  return true;
}
template <typename T> void SE3Log(const T *R9, const T *t3, T *xi) {       // :488-546
  const M3<T> R = get(R9), I = eye3<T>();
  T w[3], theta;
  M3<T> Vin = I;
  if (so3LogCore(R, w, theta)) {
    const T invTheta = T(1) / theta, invTheta2 = invTheta * invTheta;
    const M3<T> wx = skew(w[0], w[1], w[2]);
    const T A = std::sin(theta) * invTheta, B = (T(1) - std::cos(theta)) * invTheta2;
    Vin = lin(T(1), I, T(-0.5), wx, invTheta2 * (T(1) - A / (T(2) * B)), mul(wx, wx));
  }
  for (int i = 0; i < 3; ++i) xi[i] = Vin.m[i][0] * t3[0] + Vin.m[i][1] * t3[1] + Vin.m[i][2] * t3[2];
  xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
}
template <typename T> void q2r(const T *q, T *R9) {                        // :93-117
  const T qw = q[0], qx = q[1], qy = q[2], qz = q[3];
  const T qw2 = qw * qw, qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz;
  const T qxqy = qx * qy, qwqz = qw * qz, qxqz = qx * qz, qwqy = qw * qy, qwqx = qw * qx, qyqz = qy * qz;
  const T R[9] = {qw2 + qx2 - qy2 - qz2, T(2) * (qxqy - qwqz), T(2) * (qxqz + qwqy),
                  T(2) * (qxqy + qwqz), qw2 - qx2 + qy2 - qz2, T(2) * (qyqz - qwqx),
                  T(2) * (qxqz - qwqy), T(2) * (qyqz + qwqx), qw2 - qx2 - qy2 + qz2};
  for (int i = 0; i < 9; ++i) R9[i] = R[i];
}
template <typename T> void r2q(const T *R9, T *q) {                        // :206-262
  const M3<T> R = get(R9);
  const T m00 = R.m[0][0], m11 = R.m[1][1], m22 = R.m[2][2], m21 = R.m[2][1], m12 = R.m[1][2], m02 = R.m[0][2],
          m20 = R.m[2][0], m10 = R.m[1][0], m01 = R.m[0][1];
  const T tr = m00 + m11 + m22;
  T qw, qx, qy, qz;
  if (tr > 0) {
    const T S = std::sqrt(tr + T(1)) * 2; qw = T(0.25) * S; qx = (m21 - m12) / S; qy = (m02 - m20) / S; qz = (m10 - m01) / S;
  } else if ((m00 > m11) & (m00 > m22)) {
    const T S = std::sqrt(T(1) + m00 - m11 - m22) * 2; qw = (m21 - m12) / S; qx = T(0.25) * S; qy = (m01 + m10) / S; qz = (m02 + m20) / S;
  } else if (m11 > m22) {
    const T S = std::sqrt(T(1) + m11 - m00 - m22) * 2; qw = (m02 - m20) / S; qx = (m01 + m10) / S; qy = T(0.25) * S; qz = (m12 + m21) / S;
  } else {
    const T S = std::sqrt(T(1) + m22 - m00 - m11) * 2; qw = (m10 - m01) / S; qx = (m02 + m20) / S; qy = (m12 + m21) / S; qz = T(0.25) * S;
  }
  q[0] = qw; q[1] = qx; q[2] = qy; q[3] = qz;
}
template <typename T> void rotvec2q(const T *w, T *q) {                    // :146-161
  T th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < T(1e-7)) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  const T invthsinth05 = std::sin(th * T(0.5)) / th;
  q[0] = std::cos(th * T(0.5)); q[1] = w[0] * invthsinth05; q[2] = w[1] * invthsinth05; q[3] = w[2] * invthsinth05;
  const T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
template <typename T> void r2euler(const T *R9, T *e) {                    // :322-344
  const M3<T> R = get(R9);
  const T sy = std::sqrt(R.m[0][0] * R.m[0][0] + R.m[1][0] * R.m[1][0]);
  if (sy < T(1e-6)) { e[0] = std::atan2(-R.m[1][2], R.m[1][1]); e[1] = std::atan2(-R.m[2][0], sy); e[2] = 0; }
  else { e[0] = std::atan2(R.m[2][1], R.m[2][2]); e[1] = std::atan2(-R.m[2][0], sy); e[2] = std::atan2(R.m[1][0], R.m[0][0]); }
}
template <typename T> void a2r(const T *rpy, T *R9) {                      // :181-191  Rz Ry Rx
  const T r = rpy[0], p = rpy[1], y = rpy[2];
  M3<T> Rx = eye3<T>(), Ry = eye3<T>(), Rz = eye3<T>();
  Rx.m[1][1] = std::cos(r); Rx.m[1][2] = -std::sin(r); Rx.m[2][1] = std::sin(r); Rx.m[2][2] = std::cos(r);
  Ry.m[0][0] = std::cos(p); Ry.m[0][2] = std::sin(p); Ry.m[2][0] = -std::sin(p); Ry.m[2][2] = std::cos(p);
  Rz.m[0][0] = std::cos(y); Rz.m[0][1] = -std::sin(y); Rz.m[1][0] = std::sin(y); Rz.m[1][1] = std::cos(y);
  put(mul(mul(Rz, Ry), Rx), R9);
}
template <typename T> void inverseSE3(const T *R9, const T *t3, T *Ri9, T *ti3) {   // :721-736
  const M3<T> R = get(R9);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Ri9[3 * i + j] = R.m[j][i];
    ti3[i] = -(R.m[0][i] * t3[0] + R.m[1][i] * t3[1] + R.m[2][i] * t3[2]);
  }
}
template <typename T> void addFrontse3(const T *xi, const T *dxi, T *out) {  // :703-719
  T R[9], t[3], dR[9], dt[3];
  se3Exp(xi, R, t);
  se3Exp(dxi, dR, dt);
  const M3<T> Rn = mul(get(dR), get(R));
  const M3<T> D = get(dR);
  T Rn9[9], tn[3];
  put(Rn, Rn9);
  for (int i = 0; i < 3; ++i) tn[i] = D.m[i][0] * t[0] + D.m[i][1] * t[1] + D.m[i][2] * t[2] + dt[i];
  SE3Log(Rn9, tn, out);
}
template <typename T> void q1_mult_q2(const T *q1, const T *q2, T *q) {     // :74-81
  q[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
  q[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
  q[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
  q[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
}
template <typename T> int run(int op, long long n, const T *in, const T *in2, T *out) {
  static const int shape[12][3] = {{6, 0, 12}, {12, 0, 6}, {3, 0, 9}, {9, 0, 3}, {4, 0, 9}, {9, 0, 4}, {3, 0, 4},
                                   {9, 0, 3}, {3, 0, 9}, {12, 0, 12}, {6, 6, 6}, {4, 4, 4}};
  if (op < 0 || op > 11) return -1;
  const int si = shape[op][0], s2 = shape[op][1], so = shape[op][2];
  for (long long i = 0; i < n; ++i) {
    const T *a = in + i * si, *b = in2 ? in2 + i * s2 : nullptr;
    T *o = out + i * so;
    switch (op) {
      case 0: se3Exp(a, o, o + 9); break;
      case 1: SE3Log(a, a + 9, o); break;
      case 2: so3Exp(a, o); break;
      case 3: { T th; so3LogCore(get(a), o, th); break; }
      case 4: q2r(a, o); break;
      case 5: r2q(a, o); break;
      case 6: rotvec2q(a, o); break;
      case 7: r2euler(a, o); break;
      case 8: a2r(a, o); break;
      case 9: inverseSE3(a, a + 9, o, o + 9); break;
      case 10: addFrontse3(a, b, o); break;
      case 11: q1_mult_q2(a, b, o); break;
    }
  }
  return 0;
}
}  // namespace geom_ref

extern "C" {
int orc_geometry(int op, long long n, const double *in, const double *in2, double *out) { return geom_ref::run<double>(op, n, in, in2, out); }
int orc_geometry_f(int op, long long n, const float *in, const float *in2, float *out) { return geom_ref::run<float>(op, n, in, in2, out); }
}
