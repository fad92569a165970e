// GPU test driver for the pose-only drop-in C++ API (reference scene recipe: test/test_6dof_stereo_poseonly_ba.cpp:15-107
// with a seeded std::mt19937): solves one stereo 6-DoF frame and one mono 6-DoF frame, prints the estimated poses.
#include <cstdio>
#include <random>
#include <vector>

#include "core/pose_only_bundle_adjustment_solver.h"

using namespace visual_navigation::analytic_solver;

int main() {
  Eigen::Isometry3f pose_left_to_right = Eigen::Isometry3f::Identity();
  pose_left_to_right.translation() = Eigen::Vector3f(0.05f, 0.0f, 0.0f);
  Eigen::Isometry3f pose_true = Eigen::Isometry3f::Identity();
  pose_true.linear() = Eigen::AngleAxisf(-0.12f, Eigen::Vector3f::UnitY()).toRotationMatrix();
  pose_true.translation() = Eigen::Vector3f(0.4f, 0.012f, -0.5f);
  const float fx = 338.0f, fy = 338.0f, cx = 320.0f, cy = 240.0f;
  std::mt19937 gen(1234);
  std::uniform_real_distribution<float> dist_x(-1.7f, 1.7f), dist_y(-1.3f, 1.3f), dist_z(0.0f, 5.0f);
  std::vector<Eigen::Vector3f> X;
  std::vector<Eigen::Vector2f> pl, pr;
  const Eigen::Isometry3f T_cw = pose_true.inverse(), T_rl = pose_left_to_right.inverse();
  for (int i = 0; i < 1000; ++i) {
    Eigen::Vector3f w(dist_x(gen), dist_y(gen), dist_z(gen) + 1.2f);
    const Eigen::Vector3f l = T_cw * w, r = T_rl * l;
    X.push_back(w);
    pl.push_back(Eigen::Vector2f(fx * l.x() / l.z() + cx, fy * l.y() / l.z() + cy));
    pr.push_back(Eigen::Vector2f(fx * r.x() / r.z() + cx, fy * r.y() / r.z() + cy));
  }
  Options options;
  options.iteration_handle.max_num_iterations = 100;
  options.convergence_handle.threshold_cost_change = 1e-6f;
  options.convergence_handle.threshold_step_size = 1e-6f;
  options.outlier_handle.threshold_huber_loss = 1.5f;
  options.outlier_handle.threshold_outlier_rejection = 2.5f;
  PoseOnlyBundleAdjustmentSolver solver;
  int fails = 0;
  for (int stereo = 0; stereo < 2; ++stereo) {
    Eigen::Isometry3f pose = Eigen::Isometry3f::Identity();
    pose.translation() = Eigen::Vector3f(-0.2f, -0.5f, 0.0f);
    std::vector<bool> ml, mr;
    Summary summary;
    const bool ok = stereo ? solver.Solve_Stereo_6Dof(X, pl, pr, fx, fy, cx, cy, fx, fy, cx, cy, pose_left_to_right, pose, ml, mr, options, &summary)
                           : solver.Solve_Monocular_6Dof(X, pl, fx, fy, cx, cy, pose, ml, options, &summary);
    std::cout << summary.BriefReport();
    float err = 0.f;
    const Eigen::Matrix3f R = pose.linear(), Rt = pose_true.linear();
    const Eigen::Vector3f t = pose.translation(), tt = pose_true.translation();
    for (int r = 0; r < 3; ++r) { err = std::fmax(err, std::fabs(t(r) - tt(r))); for (int c = 0; c < 3; ++c) err = std::fmax(err, std::fabs(R(r, c) - Rt(r, c))); }
    std::printf("POSEONLY stereo=%d ok=%d max_err=%g debug_poses=%zu inliers=%zu\n", stereo, ok ? 1 : 0, err,
                solver.GetDebugPoses().size(), ml.size());
    if (!ok || err > 1e-3f || solver.GetDebugPoses().empty() || ml.size() != X.size()) ++fails;
  }
  std::printf(fails ? "POSEONLY_FAILED\n" : "POSEONLY_OK\n");
  return fails;
}
