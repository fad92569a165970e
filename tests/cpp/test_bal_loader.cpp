// Host-side check of ba_b200/utility/bal_loader.h: loads the file given on the command line and prints what the
// Python test compares with bundle_adjustment_solver_b200/bal.py (same conversion, independent code).
#include <cstdio>

#include "ba_b200/utility/bal_loader.h"

int main(int argc, char **argv) {
  ba_b200::BalProblem P;
  std::string err;
  if (argc < 2 || !ba_b200::LoadBal(argv[1], &P, &err)) {
    std::printf("LOAD_FAILED %s\n", err.c_str());
    return 2;
  }
  std::printf("COUNTS %zu %zu %zu\n", P.poses.size(), P.points.size(), P.obs_pose.size());
  std::printf("F0 %.17g\n", P.f0);
  for (size_t c = 0; c < P.poses.size(); ++c) {
    std::printf("POSE");
    for (int k = 0; k < 16; ++k) std::printf(" %.17g", P.poses[c].matrix().data()[k]);   // column-major 4 x 4
    std::printf("\n");
  }
  for (size_t k = 0; k < P.obs_pose.size(); ++k) std::printf("OBS %d %d %.17g %.17g\n", P.obs_pose[k], P.obs_point[k], P.obs_pixel[k](0), P.obs_pixel[k](1));
  return 0;
}
