// Drop-in for core/pose_only_bundle_adjustment_solver.h (reference :19-67): the four Solve_* entry points and
// GetDebugPoses with the reference's signatures, plus one batched entry for many independent frames.  Each
// call packs the Eigen-typed arguments into flat float buffers and runs the batched device kernel
// (ba_poseonly_solve_batched); there is no CPU path.  Link with -lba_b200.
#ifndef _POSE_ONLY_BUNDLE_ADJUSTMENT_H_
#define _POSE_ONLY_BUNDLE_ADJUSTMENT_H_

#include <chrono>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../ba_b200.h"
#include "../eigen_shim.h"
#include "solver_option_and_summary.h"

namespace visual_navigation {
namespace analytic_solver {

class PoseOnlyBundleAdjustmentSolver {
 public:
  PoseOnlyBundleAdjustmentSolver() {}
  ~PoseOnlyBundleAdjustmentSolver() {}

  const std::vector<Eigen::Isometry3f> &GetDebugPoses() const { return debug_poses_; }
  void SetDevice(int device) { device_ = device; }

  bool Solve_Monocular_Planar3Dof(const std::vector<Eigen::Vector3f> &world_position_list,
                                  const std::vector<Eigen::Vector2f> &matched_pixel_list, const float fx,
                                  const float fy, const float cx, const float cy,
                                  const Eigen::Isometry3f &pose_base_to_camera,
                                  const Eigen::Isometry3f &pose_world_to_last,
                                  Eigen::Isometry3f &pose_world_to_current, std::vector<bool> &mask_inlier,
                                  Options options, Summary *summary = nullptr) {
    CheckSizes(world_position_list.size(), matched_pixel_list.size(),
               "SolveMonocularPoseOnlyBundleAdjustment3Dof(), world_position_list.size() != current_pixel_list.size()");
    std::vector<bool> unused;
    const float intr[4] = {fx, fy, cx, cy};
    return Run(BA_POSEONLY_MONO_PLANAR3DOF, world_position_list, matched_pixel_list, nullptr, intr, intr, nullptr,
               &pose_base_to_camera, &pose_world_to_last, pose_world_to_current, mask_inlier, unused, options, summary);
  }

  bool Solve_Stereo_Planar3Dof(const std::vector<Eigen::Vector3f> &world_position_list,
                               const std::vector<Eigen::Vector2f> &matched_left_pixel_list,
                               const std::vector<Eigen::Vector2f> &matched_right_pixel_list, const float fx_left,
                               const float fy_left, const float cx_left, const float cy_left, const float fx_right,
                               const float fy_right, const float cx_right, const float cy_right,
                               const Eigen::Isometry3f &base_to_camera_pose,
                               const Eigen::Isometry3f &left_to_right_pose,
                               const Eigen::Isometry3f &world_to_last_pose, Eigen::Isometry3f &world_to_current_pose,
                               std::vector<bool> &mask_inlier_left, std::vector<bool> &mask_inlier_right,
                               Options options, Summary *summary = nullptr) {
    CheckSizes(world_position_list.size(), matched_left_pixel_list.size(),
               "SolveMonocularPoseOnlyBundleAdjustment3Dof(), world_position_list.size() != left_current_pixel_list.size()");
    CheckSizes(world_position_list.size(), matched_right_pixel_list.size(),
               "SolveMonocularPoseOnlyBundleAdjustment3Dof(), world_position_list.size() != right_current_pixel_list.size()");
    const float il[4] = {fx_left, fy_left, cx_left, cy_left}, ir[4] = {fx_right, fy_right, cx_right, cy_right};
    return Run(BA_POSEONLY_STEREO_PLANAR3DOF, world_position_list, matched_left_pixel_list, &matched_right_pixel_list, il, ir,
               &left_to_right_pose, &base_to_camera_pose, &world_to_last_pose, world_to_current_pose, mask_inlier_left,
               mask_inlier_right, options, summary);
  }

  bool Solve_Monocular_6Dof(const std::vector<Eigen::Vector3f> &reference_position_list,
                            const std::vector<Eigen::Vector2f> &matched_pixel_list, const float fx, const float fy,
                            const float cx, const float cy, Eigen::Isometry3f &reference_to_current_pose,
                            std::vector<bool> &mask_inlier, Options options, Summary *summary = nullptr) {
    CheckSizes(reference_position_list.size(), matched_pixel_list.size(),
               "SolveMonocularPoseOnlyBundleAdjustment6Dof(), world_position_list.size() != current_pixel_list.size()");
    std::vector<bool> unused;
    const float intr[4] = {fx, fy, cx, cy};
    return Run(BA_POSEONLY_MONO_6DOF, reference_position_list, matched_pixel_list, nullptr, intr, intr, nullptr, nullptr,
               nullptr, reference_to_current_pose, mask_inlier, unused, options, summary);
  }

  bool Solve_Stereo_6Dof(const std::vector<Eigen::Vector3f> &reference_position_list,
                         const std::vector<Eigen::Vector2f> &matched_left_pixel_list,
                         const std::vector<Eigen::Vector2f> &matched_right_pixel_list, const float fx_left,
                         const float fy_left, const float cx_left, const float cy_left, const float fx_right,
                         const float fy_right, const float cx_right, const float cy_right,
                         const Eigen::Isometry3f &left_to_right_pose, Eigen::Isometry3f &reference_to_current_left_pose,
                         std::vector<bool> &mask_inlier_left, std::vector<bool> &mask_inlier_right, Options options,
                         Summary *summary = nullptr) {
    CheckSizes(reference_position_list.size(), matched_left_pixel_list.size(),
               "SolveStereoPoseOnlyBundleAdjustment6Dof(), world_position_list.size() != left_current_pixel_list.size()");
    CheckSizes(reference_position_list.size(), matched_right_pixel_list.size(),
               "SolveStereoPoseOnlyBundleAdjustment6Dof(), world_position_list.size() != right_current_pixel_list.size()");
    const float il[4] = {fx_left, fy_left, cx_left, cy_left}, ir[4] = {fx_right, fy_right, cx_right, cy_right};
    return Run(BA_POSEONLY_STEREO_6DOF, reference_position_list, matched_left_pixel_list, &matched_right_pixel_list, il, ir,
               &left_to_right_pose, nullptr, nullptr, reference_to_current_left_pose, mask_inlier_left, mask_inlier_right,
               options, summary);
  }

  // Extension: many independent stereo 6-DoF frames in one device launch (BASELINE config "4096 frames x
  // ~300 observations").  frame f owns points [offsets[f], offsets[f+1]).  Returns per-frame success.
  std::vector<bool> Solve_Stereo_6Dof_Batched(const std::vector<int> &offsets, const std::vector<Eigen::Vector3f> &positions,
                                              const std::vector<Eigen::Vector2f> &left_pixels,
                                              const std::vector<Eigen::Vector2f> &right_pixels, const float fx, const float fy,
                                              const float cx, const float cy, const Eigen::Isometry3f &left_to_right_pose,
                                              std::vector<Eigen::Isometry3f> &reference_to_current_left_poses,
                                              std::vector<bool> &mask_inlier_left, std::vector<bool> &mask_inlier_right,
                                              Options options) {
    const int nf = static_cast<int>(offsets.size()) - 1;
    CheckSizes(positions.size(), left_pixels.size(), "Solve_Stereo_6Dof_Batched(), positions.size() != left_pixels.size()");
    CheckSizes(positions.size(), right_pixels.size(), "Solve_Stereo_6Dof_Batched(), positions.size() != right_pixels.size()");
    const size_t n = positions.size();
    std::vector<float> X(3 * n), pl(2 * n), pr(2 * n), poses(12 * static_cast<size_t>(nf));
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) X[3 * i + k] = positions[i](k);
      for (int k = 0; k < 2; ++k) { pl[2 * i + k] = left_pixels[i](k); pr[2 * i + k] = right_pixels[i](k); }
    }
    for (int f = 0; f < nf; ++f) Pack(reference_to_current_left_poses[f], &poses[12 * f]);
    float l2r[12];
    Pack(left_to_right_pose, l2r);
    const float intr[4] = {fx, fy, cx, cy};
    std::vector<uint8_t> ml(n), mr(n);
    std::vector<ba_poseonly_result> res(nf);
    const ba_poseonly_options o = MakeOptions(options);
    const int rc = ba_poseonly_solve_batched(device_, BA_POSEONLY_STEREO_6DOF, nf, offsets.data(), X.data(), pl.data(), pr.data(),
                                             intr, intr, l2r, nullptr, nullptr, poses.data(), ml.data(), mr.data(), &o, res.data(),
                                             nullptr, nullptr, nullptr);
    if (rc != BA_OK) throw std::runtime_error("ba_poseonly_solve_batched failed (no CUDA device? there is no CPU fallback)");
    std::vector<bool> ok(nf);
    for (int f = 0; f < nf; ++f) { ok[f] = res[f].success != 0; if (ok[f]) Unpack(&poses[12 * f], &reference_to_current_left_poses[f]); }
    mask_inlier_left.assign(ml.begin(), ml.end());
    mask_inlier_right.assign(mr.begin(), mr.end());
    return ok;
  }

 private:
  static void CheckSizes(size_t a, size_t b, const char *what) {
    if (a != b) throw std::runtime_error(std::string("In PoseOnlyBundleAdjustmentSolver::") + what);
  }
  static void Pack(const Eigen::Isometry3f &T, float *out) {
    const Eigen::Matrix3f R = T.linear();
    const Eigen::Vector3f t = T.translation();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[r * 3 + c] = R(r, c);
    for (int r = 0; r < 3; ++r) out[9 + r] = t(r);
  }
  static void Unpack(const float *in, Eigen::Isometry3f *T) {
    Eigen::Matrix3f R;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = in[r * 3 + c];
    *T = Eigen::Isometry3f::Identity();
    T->linear() = R;
    T->translation() = Eigen::Vector3f(in[9], in[10], in[11]);
  }
  static ba_poseonly_options MakeOptions(const Options &options) {
    ba_poseonly_options o;
    o.threshold_step_size = options.convergence_handle.threshold_step_size;
    o.threshold_cost_change = options.convergence_handle.threshold_cost_change;
    o.threshold_huber_loss = options.outlier_handle.threshold_huber_loss;
    o.threshold_outlier_rejection = options.outlier_handle.threshold_outlier_rejection;
    o.max_num_iterations = options.iteration_handle.max_num_iterations;
    return o;
  }
  bool Run(int kind, const std::vector<Eigen::Vector3f> &positions, const std::vector<Eigen::Vector2f> &left,
           const std::vector<Eigen::Vector2f> *right, const float *intr_l, const float *intr_r,
           const Eigen::Isometry3f *left_to_right, const Eigen::Isometry3f *base_to_camera,
           const Eigen::Isometry3f *world_to_last, Eigen::Isometry3f &pose_io, std::vector<bool> &mask_left,
           std::vector<bool> &mask_right, const Options &options, Summary *summary) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    const int max_iteration = options.iteration_handle.max_num_iterations;
    if (summary != nullptr) {
      summary->max_iteration_ = max_iteration;
      summary->threshold_cost_change_ = options.convergence_handle.threshold_cost_change;
      summary->threshold_step_size_ = options.convergence_handle.threshold_step_size;
      summary->convergence_status_ = true;
    }
    debug_poses_.resize(0);
    const size_t n = positions.size();
    std::vector<float> X(3 * n), pl(2 * n), pr(right ? 2 * n : 0);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) X[3 * i + k] = positions[i](k);
      for (int k = 0; k < 2; ++k) pl[2 * i + k] = left[i](k);
      if (right) for (int k = 0; k < 2; ++k) pr[2 * i + k] = (*right)[i](k);
    }
    float pose[12], l2r[12], b2c[12], w2l[12];
    Pack(pose_io, pose);
    if (left_to_right) Pack(*left_to_right, l2r);
    if (base_to_camera) Pack(*base_to_camera, b2c);
    if (world_to_last) Pack(*world_to_last, w2l);
    const int offsets[2] = {0, static_cast<int>(n)};
    std::vector<uint8_t> ml(n), mr(n);
    const int K = max_iteration > 0 ? max_iteration : 1;
    std::vector<float> hist_cost(K), hist_step(K), dbg(12 * static_cast<size_t>(K));
    ba_poseonly_result res{};
    const ba_poseonly_options o = MakeOptions(options);
    const int rc = ba_poseonly_solve_batched(device_, kind, 1, offsets, X.data(), pl.data(), right ? pr.data() : nullptr, intr_l,
                                             intr_r, left_to_right ? l2r : nullptr, base_to_camera ? b2c : nullptr,
                                             world_to_last ? w2l : nullptr, pose, ml.data(), mr.data(), &o, &res,
                                             hist_cost.data(), hist_step.data(), dbg.data());
    if (rc != BA_OK) throw std::runtime_error("ba_poseonly_solve_batched failed (no CUDA device? there is no CPU fallback)");
    // mask_inlier.resize(n_pts, true) then only cleared (pose_only...cpp:43,95-98)
    mask_left.assign(ml.begin(), ml.end());
    if (right) mask_right.assign(mr.begin(), mr.end());
    for (int k = 0; k < res.n_iterations && k < K; ++k) {
      Eigen::Isometry3f T;
      Unpack(&dbg[12 * k], &T);
      debug_poses_.push_back(T);
    }
    const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    if (summary != nullptr) {
      for (int k = 0; k < res.n_summary && k < K; ++k) {  // the converging trip pushes no row (:116-147)
        OptimizationInfo info;
        info.cost = hist_cost[k];
        info.cost_change = k == 0 ? std::fabs(hist_cost[0] - 1e10f) : std::fabs(hist_cost[k] - hist_cost[k - 1]);
        info.average_reprojection_error = hist_cost[k];
        info.abs_step = hist_step[k];
        info.abs_gradient = 0;
        info.damping_term = -1;
        info.iter_time = res.n_iterations > 0 ? total_ms / res.n_iterations : 0.0;
        info.iteration_status = IterationStatus::UPDATE;
        summary->optimization_info_list_.push_back(info);
      }
      summary->convergence_status_ = res.converged != 0;
      summary->total_time_in_millisecond_ = total_ms;
    }
    if (res.success) {
      Unpack(pose, &pose_io);
      return true;
    }
    std::cout << "!! WARNING !! poseonly BA yields NAN value!!\n";  // (:159-167) the pose is left untouched
    return false;
  }

  int device_{0};
  std::vector<Eigen::Isometry3f> debug_poses_;
};

}  // namespace analytic_solver
}  // namespace visual_navigation
#endif
