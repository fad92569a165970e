// Host-side check of the drop-in utility/geometry_library.h: every function of the reference's header is called
// through its reference signature; the values are printed for the Python test, which compares them with the oracle's
// restatement of utility/geometry_library.cpp.
#include <cstdio>

#include "ba_b200/utility/geometry_library.h"

template <typename M> static void dump(const char *name, const M &m, int n) {
  std::printf("%s", name);
  for (int i = 0; i < n; ++i) std::printf(" %.17g", (double)m.data()[i]);
  std::printf("\n");
}

int main() {
  geometry::Vector6d_ xi;
  const double x[6] = {0.3, -0.2, 0.5, 0.4, -0.7, 0.25};
  for (int i = 0; i < 6; ++i) xi(i) = x[i];
  Eigen::Matrix4d T;
  geometry::se3Exp(xi, T);
  dump("se3Exp", T, 16);                     // column-major 4 x 4
  geometry::Vector6d_ back;
  geometry::SE3Log(T, back);
  dump("SE3Log", back, 6);
  Eigen::Matrix3d R;
  geometry::so3Exp(Eigen::Vector3d(0.4, -0.7, 0.25), R);
  dump("so3Exp", R, 9);
  Eigen::Vector3d w;
  geometry::SO3Log(R, w);
  dump("SO3Log", w, 3);
  dump("r2q", geometry::r2q(R), 4);
  dump("q2r", geometry::q2r(geometry::r2q(R)), 9);
  dump("rotvec2q", geometry::rotvec2q(Eigen::Vector3d(0.4, -0.7, 0.25)), 4);
  dump("a2r", geometry::a2r(0.1, -0.4, 0.9), 9);
  dump("r2euler", geometry::r2euler(geometry::a2r(0.1, -0.4, 0.9)), 3);
  dump("inverseSE3", geometry::inverseSE3(T), 16);
  geometry::Vector6d_ d;
  const double dx[6] = {0.01, 0.02, -0.03, 0.05, 0.02, -0.04};
  for (int i = 0; i < 6; ++i) d(i) = dx[i];
  geometry::Vector6d_ acc = xi;
  geometry::addFrontse3(acc, d);
  dump("addFrontse3", acc, 6);
  geometry::Vector4d_ q1 = geometry::rotvec2q(Eigen::Vector3d(0.4, -0.7, 0.25)), q2 = geometry::rotvec2q(Eigen::Vector3d(-0.2, 0.1, 0.6));
  dump("q1_mult_q2", geometry::q1_mult_q2(q1, q2), 4);
  dump("q_left_mult", geometry::q_left_mult(q1), 16);
  dump("q_right_mult", geometry::q_right_mult(q1), 16);
  dump("q_conj", geometry::q_conj(q1), 4);
  dump("skewMat", geometry::skewMat(Eigen::Vector3d(1, 2, 3)), 9);
  // float variants compile and agree to float precision
  Eigen::Matrix3f Rf;
  geometry::so3Exp_f(Eigen::Vector3f(0.4f, -0.7f, 0.25f), Rf);
  dump("so3Exp_f", Rf, 9);
  return 0;
}
