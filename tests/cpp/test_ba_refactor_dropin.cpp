// GPU test driver for the refactored front-end (reference test/test_ba_refactor.cpp:235-299 call sequence) on a
// scene file written by the Python test.  usage: test_ba_refactor_dropin scene.bin result.bin max_iterations mode
// mode: lm | gn | gd (SolveByGradientDescent).  Without arguments: exception semantics on the CPU only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "core/full_bundle_adjustment_solver_refactor.h"

using namespace visual_navigation::analytic_solver;

template <typename T>
static void rd(FILE *f, T *p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { std::perror("read"); std::exit(2); } }

template <typename F>
static bool throws(F &&fn) {
  try { fn(); } catch (const std::runtime_error &) { return true; }
  return false;
}

static int registration_semantics() {
  FullBundleAdjustmentSolverRefactor s;
  OptimizerCamera cam;
  cam.fx = cam.fy = 300; cam.cx = 320; cam.cy = 240;
  cam.camera_to_body_pose = Pose::Identity();
  s.RegisterCamera(0, cam);
  s.RegisterCamera(0, cam);   // duplicate: warning, ignored
  Pose p0 = Pose::Identity(), p1 = Pose::Identity(), stranger = Pose::Identity();
  Point x0(0, 0, 5), xs(1, 1, 1);
  s.RegisterWorldToBodyPose(&p0);
  s.RegisterWorldToBodyPose(&p1);
  s.RegisterWorldPoint(&x0);
  int ok = 1;
  ok &= throws([&] { s.MakePoseFixed(nullptr); });
  ok &= throws([&] { s.MakePoseFixed(&stranger); });
  ok &= throws([&] { s.MakePointFixed(nullptr); });
  ok &= throws([&] { s.MakePointFixed(&xs); });
  ok &= throws([&] { s.AddObservation(7, &p0, &x0, Pixel(1, 2)); });
  ok &= throws([&] { s.AddObservation(0, &stranger, &x0, Pixel(1, 2)); });
  ok &= throws([&] { s.AddObservation(0, &p0, &xs, Pixel(1, 2)); });
  ok &= throws([&] { Options o; s.Solve(o); });                 // no observations
  ok &= throws([&] { Options o; s.SolveByGradientDescent(o); });
  s.MakePoseFixed(&p0);
  s.AddObservation(0, &p0, &x0, Pixel(320, 240));
  ok &= !throws([&] { s.AddObservation(0, &p1, &x0, Pixel(321, 240)); });
  std::printf(ok ? "REFACTOR_REGISTRATION_OK\n" : "REFACTOR_REGISTRATION_FAILED\n");
  return ok ? 0 : 1;
}

int main(int argc, char **argv) {
  if (argc < 5) return registration_semantics();
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[5];
  rd(f, hdr, 5);
  const int n_cam = hdr[0], n_pose = hdr[1], n_point = hdr[2], n_fixed = hdr[3], n_obs = hdr[4];
  FullBundleAdjustmentSolverRefactor ba_solver;
  for (int c = 0; c < n_cam; ++c) {
    int id; double intr[4], T[16];
    rd(f, &id, 1); rd(f, intr, 4); rd(f, T, 16);
    OptimizerCamera cam;
    cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3];
    Rotation3D R;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R(r, k) = T[k * 4 + r];
    cam.camera_to_body_pose = Pose::Identity();
    cam.camera_to_body_pose.linear() = R;
    cam.camera_to_body_pose.translation() = Translation3D(T[12], T[13], T[14]);
    ba_solver.RegisterCamera(id, cam);
  }
  std::unordered_map<int, Pose> pose_pool;
  std::unordered_map<int, Point> point_pool;
  for (int j = 0; j < n_pose; ++j) {
    double T[16];
    rd(f, T, 16);
    Pose P = Pose::Identity();
    Rotation3D R;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R(r, k) = T[k * 4 + r];
    P.linear() = R;
    P.translation() = Translation3D(T[12], T[13], T[14]);
    pose_pool[j] = P;
  }
  for (int i = 0; i < n_point; ++i) { double X[3]; rd(f, X, 3); point_pool[i] = Point(X[0], X[1], X[2]); }
  for (int j = 0; j < n_pose; ++j) ba_solver.RegisterWorldToBodyPose(&pose_pool[j]);
  for (int i = 0; i < n_point; ++i) ba_solver.RegisterWorldPoint(&point_pool[i]);
  for (int k = 0; k < n_fixed; ++k) { int j; rd(f, &j, 1); ba_solver.MakePoseFixed(&pose_pool[j]); }
  for (int k = 0; k < n_obs; ++k) {
    int ids[3]; double uv[2];
    rd(f, ids, 3); rd(f, uv, 2);
    ba_solver.AddObservation(ids[0], &pose_pool[ids[1]], &point_pool[ids[2]], Pixel(uv[0], uv[1]));
  }
  std::fclose(f);

  Options options;
  options.iteration_handle.max_num_iterations = std::atoi(argv[3]);
  options.convergence_handle.threshold_cost_change = 1e-6f;
  options.convergence_handle.threshold_step_size = 1e-6f;
  Summary summary;
  if (!std::strcmp(argv[4], "gd")) {
    ba_solver.SolveByGradientDescent(options, &summary);
  } else {
    options.solver_type = !std::strcmp(argv[4], "gn") ? SolverType::GAUSS_NEWTON : SolverType::LEVENBERG_MARQUARDT;
    ba_solver.Solve(options, &summary);
  }
  std::cout << summary.BriefReport() << std::endl;

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
