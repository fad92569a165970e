// GPU test driver for the drop-in C++ API: the same call sequence as the reference's test/test_ba.cpp
// (:235-297) on a scene read from a file (written by the Python test so that the oracle sees identical
// inputs).  usage: test_ba_dropin scene.bin result.bin [max_iterations] [accumulate]
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

#include "core/full_bundle_adjustment_solver.h"

using namespace visual_navigation::analytic_solver;

template <typename T>
static void rd(FILE *f, T *p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { std::perror("read"); std::exit(2); } }

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[5];
  rd(f, hdr, 5);
  const int n_cam = hdr[0], n_pose = hdr[1], n_point = hdr[2], n_fixed = hdr[3], n_obs = hdr[4];
  FullBundleAdjustmentSolver ba_solver;
  for (int c = 0; c < n_cam; ++c) {
    int id; double intr[4], T[16];
    rd(f, &id, 1); rd(f, intr, 4); rd(f, T, 16);
    _BA_Camera cam;
    cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3];
    _BA_Rotation3 R;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R(r, k) = T[k * 4 + r];
    cam.pose_this_to_cam0 = _BA_Pose::Identity();
    cam.pose_this_to_cam0.linear() = R;
    cam.pose_this_to_cam0.translation() = _BA_Position3(T[12], T[13], T[14]);
    ba_solver.AddCamera(id, cam);
  }
  // the caller owns the parameters; the solver keys them by address (unordered_map nodes are stable)
  std::unordered_map<int, _BA_Pose> pose_pool;
  std::unordered_map<int, _BA_Point> point_pool;
  for (int j = 0; j < n_pose; ++j) {
    double T[16];
    rd(f, T, 16);
    _BA_Pose P = _BA_Pose::Identity();
    _BA_Rotation3 R;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R(r, k) = T[k * 4 + r];
    P.linear() = R;
    P.translation() = _BA_Position3(T[12], T[13], T[14]);
    pose_pool[j] = P;
  }
  for (int i = 0; i < n_point; ++i) { double X[3]; rd(f, X, 3); point_pool[i] = _BA_Point(X[0], X[1], X[2]); }
  for (int j = 0; j < n_pose; ++j) ba_solver.AddPose(&pose_pool[j]);
  for (int i = 0; i < n_point; ++i) ba_solver.AddPoint(&point_pool[i]);
  for (int k = 0; k < n_fixed; ++k) { int j; rd(f, &j, 1); ba_solver.MakePoseFixed(&pose_pool[j]); }
  ba_solver.MakePointFixed({});
  for (int k = 0; k < n_obs; ++k) {
    int ids[3]; double uv[2];
    rd(f, ids, 3); rd(f, uv, 2);
    ba_solver.AddObservation(ids[0], &pose_pool[ids[1]], &point_pool[ids[2]], _BA_Pixel(uv[0], uv[1]));
  }
  std::fclose(f);

  Options options;
  options.iteration_handle.max_num_iterations = argc > 3 ? std::atoi(argv[3]) : 3000;
  options.convergence_handle.threshold_cost_change = 1e-6f;
  options.convergence_handle.threshold_step_size = 1e-6f;
  options.accumulate_offdiagonal_blocks = argc > 4 && std::atoi(argv[4]) != 0;
  Summary summary;
  ba_solver.Solve(options, &summary);
  std::cout << summary.BriefReport() << std::endl;
  const std::string full = summary.FullReport();   // declared by the reference (:83), defined only here
  if (full.find("Analytic Solver Full Report") == std::string::npos || full.find("Stopped because") == std::string::npos ||
      full.size() <= summary.BriefReport().size()) {
    std::cout << "FULL_REPORT_BAD" << std::endl;
    return 3;
  }
  std::cout << full.substr(summary.BriefReport().size()) << std::endl;

  FILE *o = std::fopen(argv[2], "wb");
  const auto &infos = summary.optimization_info_list();
  const int n_it = static_cast<int>(infos.size()), conv = summary.convergence_status() ? 1 : 0;
  std::fwrite(&n_it, sizeof(int), 1, o);
  std::fwrite(&conv, sizeof(int), 1, o);
  for (const auto &i : infos) { std::fwrite(&i.cost, sizeof(double), 1, o); }
  for (int j = 0; j < n_pose; ++j) std::fwrite(pose_pool[j].data(), sizeof(double), 16, o);
  for (int i = 0; i < n_point; ++i) std::fwrite(point_pool[i].data(), sizeof(double), 3, o);
  std::fclose(o);
  return 0;
}
