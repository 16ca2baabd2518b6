// TEST INFRASTRUCTURE — C entry points around the UNMODIFIED reference P3P (compiled from /root/reference by oracle/Makefile
// into oracle/_ref/libref_p3p.so).  Same calling convention as mpeo_p3p / mpeo_solve_quartic in pose_oracle.cpp.
#include "monocular_pose_estimator_lib/p3p.h"

extern "C" {

int ref_p3p(const double f[9], const double P[9], double sol[48]) {
  Eigen::Matrix3d fv, wp;
  for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) { fv(r, k) = f[3 * k + r]; wp(r, k) = P[3 * k + r]; }
  Eigen::Matrix<Eigen::Matrix<double, 3, 4>, 4, 1> s;
  for (int i = 0; i < 48; ++i) sol[i] = 0;
  int rc = monocular_pose_estimator::P3P::computePoses(fv, wp, s);
  if (rc == 0)
    for (int i = 0; i < 4; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) sol[12 * i + 4 * r + c] = s(i)(r, c);
  return rc;
}

int ref_solve_quartic(const double factors[5], double roots[4]) {
  Eigen::Matrix<double, 5, 1> f; Eigen::Matrix<double, 4, 1> r;
  for (int i = 0; i < 5; ++i) f(i) = factors[i];
  int rc = monocular_pose_estimator::P3P::solveQuartic(f, r);
  for (int i = 0; i < 4; ++i) roots[i] = r(i);
  return rc;
}

}
