// CPU test of the drop-in class's registration semantics (no GPU needed: nothing here calls Solve).
// Mirrors the behaviours of core/full_bundle_adjustment_solver.cpp:87-180.
#include <cstdio>
#include <stdexcept>

#include "core/full_bundle_adjustment_solver.h"
#include "core/pose_only_bundle_adjustment_solver.h"

using namespace visual_navigation::analytic_solver;

#define EXPECT(cond)                                                    \
  do {                                                                  \
    if (!(cond)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main() {
  static_assert(sizeof(_BA_Pose) == 128, "Isometry3d is a column-major 4x4");
  static_assert(sizeof(_BA_Point) == 24 && sizeof(_BA_Pixel) == 16, "fixed-size vectors");
  FullBundleAdjustmentSolver solver;
  _BA_Camera cam;
  cam.fx = cam.fy = 500.0; cam.cx = 200.0; cam.cy = 100.0;
  cam.pose_this_to_cam0 = _BA_Pose::Identity();
  solver.AddCamera(0, cam);
  solver.AddCamera(0, cam);  // duplicate id ignored
  _BA_Pose p0 = _BA_Pose::Identity(), p1 = _BA_Pose::Identity(), stranger = _BA_Pose::Identity();
  p1.translation() = _BA_Position3(0.3, 0.0, 0.0);
  _BA_Point x0(0.0, 0.0, 5.0), x1(1.0, 0.5, 4.0), lonely(0, 0, 1);
  solver.AddPose(&p0); solver.AddPose(&p1); solver.AddPose(&p1);  // duplicate pointer ignored
  solver.AddPoint(&x0); solver.AddPoint(&x1);
  solver.MakePoseFixed(nullptr);   // message, no throw (test_ba.cpp:252 passes {} deliberately)
  solver.MakePointFixed({});
  bool threw = false;
  try { solver.MakePoseFixed(&stranger); } catch (const std::runtime_error &) { threw = true; }
  EXPECT(threw);
  threw = false;
  try { solver.MakePointFixed(&lonely); } catch (const std::runtime_error &) { threw = true; }
  EXPECT(threw);
  solver.MakePoseFixed(&p0);
  solver.AddObservation(7, &p0, &x0, _BA_Pixel(1, 2));        // invalid camera: dropped
  solver.AddObservation(0, &stranger, &x0, _BA_Pixel(1, 2));  // unknown pose: dropped
  solver.AddObservation(0, &p0, &lonely, _BA_Pixel(1, 2));    // unknown point: dropped
  solver.AddObservation(0, &p0, &x0, _BA_Pixel(200, 100));
  solver.FinalizeParameters();
  solver.FinalizeParameters();     // idempotent
  _BA_Pose late = _BA_Pose::Identity();
  solver.AddPose(&late);           // after finalize: warning, ignored
  solver.AddObservation(0, &p1, &x1, _BA_Pixel(10, 20));  // still accepted after finalize
  solver.GetSolverStatistics();
  // isometry inverse used by AddPose
  _BA_Pose T = _BA_Pose::Identity();
  T.linear() = Eigen::AngleAxis<double>(0.3, _BA_Position3::UnitZ()).toRotationMatrix();
  T.translation() = _BA_Position3(1, 2, 3);
  const _BA_Pose I = T * T.inverse();
  for (int r = 0; r < 3; ++r) {
    const _BA_Rotation3 R = I.linear();
    const _BA_Position3 t = I.translation();
    EXPECT(std::fabs(R(r, r) - 1.0) < 1e-14 && std::fabs(t(r)) < 1e-14);
  }
  // pose-only: size mismatch throws like the reference (pose_only...cpp:31-37)
  PoseOnlyBundleAdjustmentSolver po;
  std::vector<Eigen::Vector3f> X(3);
  std::vector<Eigen::Vector2f> px(2);
  std::vector<bool> mask;
  Eigen::Isometry3f pose = Eigen::Isometry3f::Identity();
  threw = false;
  try { po.Solve_Monocular_6Dof(X, px, 300, 300, 320, 240, pose, mask, Options()); } catch (const std::runtime_error &) { threw = true; }
  EXPECT(threw);
  Summary s;
  s.BriefReport();  // empty summary must not crash
  s.FullReport();
  std::printf("REGISTRATION_OK\n");
  return 0;
}
