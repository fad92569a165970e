// bal_loader.h -- reader of the "Bundle Adjustment in the Large" text format for the drop-in solver (the on-disk
// form of the Venice data set BASELINE config C5 is shaped after; the reference has no on-disk format, SURVEY.md 8f
// rank 4).  Same conversion as bundle_adjustment_solver_b200/bal.py, whose docstring derives it:
//   file model   P = R X + t, p = -P / P.z, pixel = f (1 + k1 |p|^2 + k2 |p|^4) p
//   solver model ideal pinhole looking down +z, one rig camera shared by all poses (full_bundle_adjustment_solver.cpp:744-760)
// so every camera frame is turned by pi about x (D = diag(1,-1,-1)), v changes sign, the radial distortion is
// removed from the pixels and camera c's pixels are rescaled by f0 / f_c (f0 = median focal length).
#ifndef BA_B200_BAL_LOADER_H_
#define BA_B200_BAL_LOADER_H_

#include <algorithm>
#include <cmath>
#include <fstream>
#include <string>
#include <vector>

#include "../eigen_shim.h"
#include "geometry_math.h"

namespace ba_b200 {

struct BalProblem {
  typedef Eigen::Isometry3d Pose;                // camera-to-world: what AddPose receives
  std::vector<Pose> poses;
  std::vector<Eigen::Vector3d> points;
  std::vector<int> obs_pose, obs_point;
  std::vector<Eigen::Vector2d> obs_pixel;        // undistorted, in units of the rig camera (focal f0, principal point 0)
  std::vector<double> focal, k1, k2;             // as in the file
  double f0 = 0.0;
};

inline bool LoadBal(const std::string &path, BalProblem *out, std::string *error = nullptr) {
  auto fail = [&](const char *msg) { if (error) *error = path + ": " + msg; return false; };
  std::ifstream in(path);
  if (!in) return fail("cannot open");
  long long nc = 0, np = 0, no = 0;
  if (!(in >> nc >> np >> no) || nc <= 0 || np <= 0 || no < 0) return fail("bad header");
  BalProblem &P = *out;
  P = BalProblem();
  P.obs_pose.resize(no); P.obs_point.resize(no); P.obs_pixel.resize(no);
  std::vector<double> px(2 * no);
  for (long long k = 0; k < no; ++k) {
    if (!(in >> P.obs_pose[k] >> P.obs_point[k] >> px[2 * k] >> px[2 * k + 1])) return fail("truncated observation list");
    if (P.obs_pose[k] < 0 || P.obs_pose[k] >= nc || P.obs_point[k] < 0 || P.obs_point[k] >= np) return fail("observation index out of range");
  }
  std::vector<double> cam(9 * nc);
  for (double &v : cam) if (!(in >> v)) return fail("truncated camera block");
  P.points.resize(np);
  for (long long i = 0; i < np; ++i) {
    double x, y, z;
    if (!(in >> x >> y >> z)) return fail("truncated point block");
    P.points[i] = Eigen::Vector3d(x, y, z);
  }
  P.focal.resize(nc); P.k1.resize(nc); P.k2.resize(nc);
  P.poses.resize(nc);
  for (long long c = 0; c < nc; ++c) {
    const double *q = &cam[9 * c];
    P.focal[c] = q[6]; P.k1[c] = q[7]; P.k2[c] = q[8];
    double R[9], Ri[9], ti[3];
    ba_geom::so3_exp(q, R);
    const double t[3] = {q[3], -q[4], -q[5]};                 // D t
    for (int k = 3; k < 9; ++k) R[k] = -R[k];                 // D R: rows 1 and 2 change sign
    ba_geom::inverse_se3(R, t, Ri, ti);                       // camera-to-world
    BalProblem::Pose T;
    T.setIdentity();
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 3; ++k) T.linear()(r, k) = Ri[3 * r + k];
      T.translation()(r) = ti[r];
    }
    P.poses[c] = T;
  }
  std::vector<double> fs = P.focal;
  std::nth_element(fs.begin(), fs.begin() + fs.size() / 2, fs.end());
  P.f0 = fs[fs.size() / 2];
  if (fs.size() % 2 == 0) P.f0 = 0.5 * (P.f0 + *std::max_element(fs.begin(), fs.begin() + fs.size() / 2));
  for (long long k = 0; k < no; ++k) {
    const int c = P.obs_pose[k];
    const double dx = px[2 * k] / P.focal[c], dy = px[2 * k + 1] / P.focal[c];
    double x = dx, y = dy;
    for (int it = 0; it < 20; ++it) {                         // fixed-point inversion of the radial polynomial
      const double r2 = x * x + y * y, s = 1.0 + P.k1[c] * r2 + P.k2[c] * r2 * r2;
      x = dx / s; y = dy / s;
    }
    P.obs_pixel[k] = Eigen::Vector2d(P.f0 * x, -P.f0 * y);
  }
  return true;
}

// Registers the problem with a drop-in FullBundleAdjustmentSolver (or the refactored front-end through its own
// adapter): one rig camera (id 0), every pose and point by address -- `problem` must outlive the solve.
template <typename Solver, typename Camera>
inline void RegisterBal(Solver &solver, BalProblem &problem, int num_fixed_poses = 2) {
  Camera cam;
  cam.fx = problem.f0; cam.fy = problem.f0; cam.cx = 0.0; cam.cy = 0.0;
  cam.pose_this_to_cam0.setIdentity();
  solver.AddCamera(0, cam);
  for (auto &T : problem.poses) solver.AddPose(&T);
  for (auto &X : problem.points) solver.AddPoint(&X);
  for (int j = 0; j < num_fixed_poses && j < (int)problem.poses.size(); ++j) solver.MakePoseFixed(&problem.poses[j]);
  for (size_t k = 0; k < problem.obs_pose.size(); ++k)
    solver.AddObservation(0, &problem.poses[problem.obs_pose[k]], &problem.points[problem.obs_point[k]], problem.obs_pixel[k]);
}

}  // namespace ba_b200
#endif  // BA_B200_BAL_LOADER_H_
