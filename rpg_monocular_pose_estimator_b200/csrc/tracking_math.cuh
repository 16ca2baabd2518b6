// The small sequential helpers of the tracking branch of PoseEstimator::estimateBodyPose, written once for the device
// (k4_tracking.cu: one thread per stream) and for the host (the mpe_host_* entry points of mpe_abi.cu, which the stage-by-stage
// mirrors — the C++ shim and the Python class — call, so that there is ONE restatement of this arithmetic in the product):
// predictPose (pose_estimator.cpp:232-244), logarithmMap (:996-1064), exponentialMap (:962-994), project2d (:251-268),
// LEDDetector::distortPoints / determineROI (led_detector.cpp:181-224, :114-179).  Operation order = oracle/pose_oracle.cpp.
#pragma once
#include <math.h>
#include "mpe_internal.cuh"

#if defined(__CUDACC__)
#define MPE_TM __host__ __device__ inline
#else
#define MPE_TM inline
#endif

namespace mpe {

struct M4 { double m[16]; };   // row-major

MPE_TM M4 m4_mul(const M4& A, const M4& B) {
  M4 C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = A.m[4 * i] * B.m[j];
      s += A.m[4 * i + 1] * B.m[4 + j];
      s += A.m[4 * i + 2] * B.m[8 + j];
      s += A.m[4 * i + 3] * B.m[12 + j];
      C.m[4 * i + j] = s;
    }
  return C;
}

// General 4x4 inverse by cofactors — the formula of oracle/pose_oracle.cpp inverse4 (previous_pose_.inverse(), :235)
MPE_TM M4 m4_inverse(const M4& A) {
  const double* a = A.m;
  double inv[16];
  inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
  inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
  inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
  inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
  inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
  inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
  inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
  inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
  inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
  inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
  inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
  inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
  inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
  double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
  double idet = 1.0 / det;
  M4 R;
  for (int i = 0; i < 16; ++i) R.m[i] = inv[i] * idet;
  return R;
}

// pose_estimator.cpp:962-994
MPE_TM M4 exponential_map(const double twist[6]) {
  const double ux = twist[0], uy = twist[1], uz = twist[2], wx = twist[3], wy = twist[4], wz = twist[5];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz), theta_squared = theta * theta;
  const double O[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
  double O2[3][3], rot[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  if (theta == 0) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    const double s = sin(theta), c = cos(theta);
    const double kv1 = (1 - c) / (theta_squared), kv2 = (theta - s) / (theta_squared * theta);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double I = (i == j) ? 1.0 : 0.0;
        rot[i][j] = I + O[i][j] / theta * s + O2[i][j] / theta_squared * (1 - c);
        V[i][j] = I + kv1 * O[i][j] + kv2 * O2[i][j];
      }
  }
  M4 T;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T.m[4 * r + c] = rot[r][c];
    T.m[4 * r + 3] = V[r][0] * ux + V[r][1] * uy + V[r][2] * uz;
  }
  T.m[12] = 0; T.m[13] = 0; T.m[14] = 0; T.m[15] = 1;
  return T;
}

// pose_estimator.cpp:996-1064 (same special cases as the oracle's restatement)
MPE_TM void logarithm_map(const M4& trans, double xi[6]) {
  double R[3][3], t[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r][c] = trans.m[4 * r + c]; t[r] = trans.m[4 * r + 3]; }
  double w_hat[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double dn = 0, rn = 0;
  for (int j = 0; j < 3; ++j)      // column-major, the order Eigen's squaredNorm() visits a Matrix3d
    for (int i = 0; i < 3; ++i) { const double I = (i == j) ? 1.0 : 0.0; dn += (R[i][j] - I) * (R[i][j] - I); rn += R[i][j] * R[i][j]; }
  const bool approx_identity = dn <= 1e-10 * 1e-10 * fmin(rn, 3.0);     // R.isApprox(I, 1e-10)
  if (!approx_identity) {
    double temp = (R[0][0] + R[1][1] + R[2][2] - 1) / 2;
    if (temp > 1) temp = 1; else if (temp < -1) temp = -1;
    const double phi = acos(temp);
    if (phi != 0) {
      const double s2 = 2 * sin(phi);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w_hat[i][j] = (R[i][j] - R[j][i]) / s2 * phi;
    }
  }
  const double w[3] = {w_hat[2][1], w_hat[0][2], w_hat[1][0]};
  const double w_norm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double A_inv[3][3];
  const bool t_zero = (t[0] == 0 && t[1] == 0 && t[2] == 0);           // t.isApproxToConstant(0, 1e-10)
  if (t_zero) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A_inv[i][j] = 0;
  } else if (w_norm == 0 || sin(w_norm) == 0) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A_inv[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    const double k = (2 * sin(w_norm) - w_norm * (1 + cos(w_norm))) / (2 * w_norm * w_norm * sin(w_norm));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        // `I - w_hat / 2 + k * w_hat * w_hat` parses as (I - w_hat/2) + ((k * w_hat) * w_hat)   (:1056-1057)
        const double w2 = (k * w_hat[i][0]) * w_hat[0][j] + (k * w_hat[i][1]) * w_hat[1][j] + (k * w_hat[i][2]) * w_hat[2][j];
        A_inv[i][j] = (((i == j) ? 1.0 : 0.0) - w_hat[i][j] / 2) + w2;
      }
  }
  for (int r = 0; r < 3; ++r) xi[r] = A_inv[r][0] * t[0] + A_inv[r][1] * t[1] + A_inv[r][2] * t[2];
  xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
}

// project2d (pose_estimator.cpp:251-268): (K|0) * T first, then * p, then divide by z
MPE_TM void project2d(const double K[9], const M4& T, double x, double y, double z, double& u, double& v) {
  double KT[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) KT[4 * i + j] = K[3 * i] * T.m[j] + K[3 * i + 1] * T.m[4 + j] + K[3 * i + 2] * T.m[8 + j];
  const double t0 = KT[0] * x + KT[1] * y + KT[2] * z + KT[3] * 1.0;
  const double t1 = KT[4] * x + KT[5] * y + KT[6] * z + KT[7] * 1.0;
  const double t2 = KT[8] * x + KT[9] * y + KT[10] * z + KT[11] * 1.0;
  u = t0 / t2;
  v = t1 / t2;
}

// LEDDetector::distortPoints for one point (led_detector.cpp:181-224): float in, float out, double arithmetic
MPE_TM void distort_point(const DevCamera& cam, float sx, float sy, float& ox, float& oy) {
  const double fx = cam.K[0], fy = cam.K[4], cx = cam.K[2], cy = cam.K[5];
  const double k1 = cam.D[0], k2 = cam.D[1], p1 = cam.D[2], p2 = cam.D[3], k3 = cam.D[4];
  const double x = ((double)sx - cx) / fx, y = ((double)sy - cy) / fy;
  const double r2 = x * x + y * y;
  double xc = x * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2);
  double yc = y * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2);
  xc = xc + (2. * p1 * x * y + p2 * (r2 + 2. * x * x));
  yc = yc + (p1 * (r2 + 2. * y * y) + 2. * p2 * x * y);
  xc = xc * fx + cx;
  yc = yc * fy + cy;
  ox = (float)xc;
  oy = (float)yc;
}

// LEDDetector::determineROI (led_detector.cpp:114-179) on predicted pixel positions px[n][2]; whole image when degenerate
MPE_TM Roi determine_roi(const DevCamera& cam, const double* px, int n, int img_w, int img_h, int roi_border) {
  Roi roi; roi.x = 0; roi.y = 0; roi.w = img_w; roi.h = img_h;
  double x_min = HUGE_VAL, x_max = 0, y_min = HUGE_VAL, y_max = 0;
  for (int i = 0; i < n; ++i) {
    const double u = px[2 * i], v = px[2 * i + 1];
    if (u < x_min) x_min = u;
    if (u > x_max) x_max = u;
    if (v < y_min) y_min = v;
    if (v > y_max) y_max = v;
  }
  float dax, day, dbx, dby;
  distort_point(cam, (float)x_min, (float)y_min, dax, day);      // the corners go through Point2f (:144-145)
  distort_point(cam, (float)x_max, (float)y_max, dbx, dby);
  const double border = (double)roi_border;
  const double x0 = fmax(0.0, fmin((double)img_w, (double)dax - border));
  const double x1 = fmax(0.0, fmin((double)img_w, (double)dbx + border));
  const double y0 = fmax(0.0, fmin((double)img_h, (double)day - border));
  const double y1 = fmax(0.0, fmin((double)img_h, (double)dby + border));
  if (!(x1 - x0 < 1 || y1 - y0 < 1) && (x1 - x0 == x1 - x0) && (y1 - y0 == y1 - y0)) {
    roi.x = (int)x0; roi.y = (int)y0; roi.w = (int)(x1 - x0); roi.h = (int)(y1 - y0);
  }
  return roi;
}

// predictPose (pose_estimator.cpp:232-244)
MPE_TM M4 predict_pose(const M4& prev, const M4& cur, double previous_time, double current_time, double predicted_time) {
  double delta[6], delta_hat[6];
  logarithm_map(m4_mul(m4_inverse(prev), cur), delta);
  for (int i = 0; i < 6; ++i) delta_hat[i] = delta[i] / (current_time - previous_time) * (predicted_time - current_time);
  return m4_mul(cur, exponential_map(delta_hat));
}

}  // namespace mpe
