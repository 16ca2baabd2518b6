// K4 — the tracking branch of PoseEstimator::estimateBodyPose with device-resident per-stream state
// (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:97-144 and the helpers it calls: predictWithROI :814-829,
// predictPose :232-244, logarithmMap :996-1064, exponentialMap :962-994, predictMarkerPositionsInImage :270-276,
// LEDDetector::determineROI led_detector.cpp:114-179 with distortPoints :181-224, findCorrespondences :372-392,
// findCorrespondencesAndPredictPose :831-848, optimiseAndUpdatePose :802-812, updatePose :794-800).
//
// One thread per stream: these steps are a few hundred sequential flops each; what matters is that thousands of
// independent streams advance together without a host round trip.  The heavy stages in between (K1 on the predicted ROI,
// K3 check / Gauss-Newton, K2 only for streams that must re-initialise) are the same kernels as in cold mode, steered by
// per-stream activity masks that these small kernels write.
#include "mpe_internal.cuh"

namespace mpe {

namespace {

struct M4 { double m[16]; };   // row-major

__device__ __forceinline__ M4 m4_mul(const M4& A, const M4& B) {
  M4 C;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
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
__device__ M4 m4_inverse(const M4& A) {
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
#pragma unroll
  for (int i = 0; i < 16; ++i) R.m[i] = inv[i] * idet;
  return R;
}

// pose_estimator.cpp:962-994
__device__ M4 exponential_map(const double twist[6]) {
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
__device__ void logarithm_map(const M4& trans, double xi[6]) {
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
__device__ __forceinline__ void project2d(const double K[9], const M4& T, double x, double y, double z, double& u, double& v) {
  double KT[12];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) KT[4 * i + j] = K[3 * i] * T.m[j] + K[3 * i + 1] * T.m[4 + j] + K[3 * i + 2] * T.m[8 + j];
  const double t0 = KT[0] * x + KT[1] * y + KT[2] * z + KT[3] * 1.0;
  const double t1 = KT[4] * x + KT[5] * y + KT[6] * z + KT[7] * 1.0;
  const double t2 = KT[8] * x + KT[9] * y + KT[10] * z + KT[11] * 1.0;
  u = t0 / t2;
  v = t1 / t2;
}

// LEDDetector::distortPoints for one point (led_detector.cpp:181-224): float in, float out, double arithmetic
__device__ __forceinline__ void distort_point(const DevCamera& cam, float sx, float sy, float& ox, float& oy) {
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

}  // namespace

// ---- step 1: predictWithROI (or the cold-branch preamble) ---------------------------------------------------------
__global__ void track_begin_kernel(const TrackArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  StreamState& st = a.state[s];
  const double t = a.times[s];
  a.a_retry[s] = 0; a.a_check[s] = 0; a.a_init[s] = 0; a.a_gn[s] = 0; a.done[s] = 0;
  a.ok[s] = 0; a.iters[s] = 0; a.updated[s] = 0; a.n_corr[s] = 0; a.track_flags[s] = 0;
  Roi roi; roi.x = 0; roi.y = 0; roi.w = a.img_w; roi.h = a.img_h;
  if (st.it_since_initialized < 1) {                         // cold branch (:68-96)
    st.predicted_time = t;
    a.mode[s] = 0;
  } else {                                                   // predictWithROI (:814-829)
    a.mode[s] = 1;
    M4 cur, prev, pred;
    for (int i = 0; i < 16; ++i) { cur.m[i] = st.current_pose[i]; prev.m[i] = st.previous_pose[i]; pred.m[i] = st.predicted_pose[i]; }
    if (st.it_since_initialized >= 2) {                      // predictPose (:232-244)
      st.predicted_time = t;
      double delta[6], delta_hat[6];
      logarithm_map(m4_mul(m4_inverse(prev), cur), delta);
      for (int i = 0; i < 6; ++i) delta_hat[i] = delta[i] / (st.current_time - st.previous_time) * (st.predicted_time - st.current_time);
      pred = m4_mul(cur, exponential_map(delta_hat));
      for (int i = 0; i < 16; ++i) st.predicted_pose[i] = pred.m[i];
    } else {
      st.predicted_time = t;
    }
    // predictMarkerPositionsInImage (:270-276) + determineROI (led_detector.cpp:114-179)
    double x_min = HUGE_VAL, x_max = 0, y_min = HUGE_VAL, y_max = 0;
    for (int i = 0; i < a.pp.n_obj; ++i) {
      double u, v;
      project2d(a.cam.K, pred, a.pp.markers[3 * i], a.pp.markers[3 * i + 1], a.pp.markers[3 * i + 2], u, v);
      a.pred_px[((size_t)s * MPE_MAX_LEDS + i) * 2] = u;
      a.pred_px[((size_t)s * MPE_MAX_LEDS + i) * 2 + 1] = v;
      if (u < x_min) x_min = u;
      if (u > x_max) x_max = u;
      if (v < y_min) y_min = v;
      if (v > y_max) y_max = v;
    }
    float dax, day, dbx, dby;
    distort_point(a.cam, (float)x_min, (float)y_min, dax, day);      // the corners go through Point2f (:144-145)
    distort_point(a.cam, (float)x_max, (float)y_max, dbx, dby);
    const double border = (double)a.roi_border;
    const double x0 = fmax(0.0, fmin((double)a.img_w, (double)dax - border));
    const double x1 = fmax(0.0, fmin((double)a.img_w, (double)dbx + border));
    const double y0 = fmax(0.0, fmin((double)a.img_h, (double)day - border));
    const double y1 = fmax(0.0, fmin((double)a.img_h, (double)dby + border));
    if (!(x1 - x0 < 1 || y1 - y0 < 1) && (x1 - x0 == x1 - x0) && (y1 - y0 == y1 - y0)) {
      roi.x = (int)x0; roi.y = (int)y0; roi.w = (int)(x1 - x0); roi.h = (int)(y1 - y0);
    }
  }
  a.rois[s] = roi;
  a.result_rois[s] = roi;
  for (int i = 0; i < 16; ++i) a.pose_io[(size_t)s * 16 + i] = st.predicted_pose[i];
}

// ---- step 2 (pass 0) / 2' (pass 1): what to do with the detections -----------------------------------------------
__device__ __forceinline__ void track_after_detect_body(const TrackArgs& a, int pass, int s) {
  if (a.done[s]) return;
  if (pass == 1 && !a.a_retry[s]) return;
  const int n = a.n_det[s];
  const bool enough = (n >= 4 && n <= MPE_MAX_DET);                  // min_num_leds_detected_ = 4 (pose_estimator.h:78)
  if (n > MPE_MAX_DET) a.track_flags[s] |= MPE_F_TOO_MANY_DET;
  if (pass == 1) {
    a.a_retry[s] = 0;
    Roi none; none.x = 0; none.y = 0; none.w = 0; none.h = 0;
    a.rois[s] = none;
  }
  if (a.mode[s] == 0) {                                              // cold: initialise() if enough LEDs (:80-91)
    if (enough) a.a_init[s] = 1; else a.done[s] = 1;
    return;
  }
  if (enough) {                                                      // findCorrespondences (:372-392, :862-906)
    const double* det = a.det + (size_t)s * MPE_MAX_BLOBS * 2;
    uint32_t* corr = a.corr + (size_t)s * 2 * MPE_MAX_LEDS;
    int k = 0;
    for (int i = 0; i < a.pp.n_obj; ++i) {
      const double pu = a.pred_px[((size_t)s * MPE_MAX_LEDS + i) * 2], pv = a.pred_px[((size_t)s * MPE_MAX_LEDS + i) * 2 + 1];
      double best = HUGE_VAL;
      uint32_t bj = 0;
      for (int j = 0; j < n; ++j) {
        const double dx = pu - det[2 * j], dy = pv - det[2 * j + 1];
        const double d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; bj = (uint32_t)j + 1; }
      }
      if (sqrt(best) <= a.pp.nearest_neighbour_pixel_tolerance) { corr[2 * k] = (uint32_t)i + 1; corr[2 * k + 1] = bj; ++k; }
    }
    a.n_corr[s] = k;
    a.a_check[s] = 1;
  } else if (pass == 0) {                                            // too few LEDs in the ROI: search the whole image once (:122-134)
    a.a_retry[s] = 1;
    Roi full; full.x = 0; full.y = 0; full.w = a.img_w; full.h = a.img_h;
    a.rois[s] = full;
    a.result_rois[s] = full;
    a.track_flags[s] |= MPE_F_FULL_IMAGE_RETRY;
  } else {
    a.done[s] = 1;
  }
}

__global__ void track_after_detect_kernel(const TrackArgs a, int pass) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  track_after_detect_body(a, pass, s);
}

// pass 0 wrapper: afterwards the ROI of every stream that does not retry is emptied, so that the retry launches of K1 find no
// tile to work on for it
__global__ void track_after_detect0_kernel(const TrackArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  track_after_detect_body(a, 0, s);
  if (!a.a_retry[s]) { Roi none; none.x = 0; none.y = 0; none.w = 0; none.h = 0; a.rois[s] = none; }
}

// (steps 3 and 4 — "checkCorrespondences succeeded -> optimisePose, else initialise()" (:835-846) and "initialise() succeeded ->
//  optimisePose" — are written by the check kernel's epilogue: K3Args::set_gn_if_ok / set_init_if_fail.)

// ---- step 5: optimiseAndUpdatePose bookkeeping (:802-812, :794-800) + result records ------------------------------
__global__ void track_finish_kernel(const TrackArgs a, mpe_result* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  StreamState& st = a.state[s];
  const bool upd = a.a_gn[s] && a.updated[s];
  if (upd) {
    for (int i = 0; i < 16; ++i) st.predicted_pose[i] = a.pose_io[(size_t)s * 16 + i];
    if (st.it_since_initialized < 2) st.it_since_initialized++;
    for (int i = 0; i < 16; ++i) { st.previous_pose[i] = st.current_pose[i]; st.current_pose[i] = st.predicted_pose[i]; }
    st.previous_time = st.current_time;
    st.current_time = st.predicted_time;
  } else if (a.ok[s]) {
    for (int i = 0; i < 16; ++i) st.predicted_pose[i] = a.pose_io[(size_t)s * 16 + i];
  }
  mpe_result r;
  r.updated = upd ? 1 : 0;
  r.n_det = a.n_det[s];
  r.n_corr = a.n_corr[s];
  r.gn_iters = a.iters[s];
  r.flags = a.flags[s] | a.track_flags[s] | (a.a_init[s] ? MPE_F_INITIALISED : 0);
  r.init_ok = a.ok[s];
  const Roi roi = a.result_rois[s];
  r.roi.x = roi.x; r.roi.y = roi.y; r.roi.width = roi.w; r.roi.height = roi.h;
  for (int i = 0; i < 16; ++i) r.pose[i] = st.predicted_pose[i];
  for (int i = 0; i < 36; ++i) r.cov[i] = upd ? a.cov[(size_t)s * 36 + i] : 0.0;
  for (int i = 0; i < 2 * MPE_MAX_LEDS; ++i) r.corr[i] = (i < 2 * r.n_corr) ? a.corr[(size_t)s * 2 * MPE_MAX_LEDS + i] : 0u;
  const int nd = r.n_det < MPE_MAX_DET ? r.n_det : MPE_MAX_DET;
  for (int i = 0; i < 2 * MPE_MAX_DET; ++i) {
    r.det[i] = (i < 2 * nd) ? a.det[(size_t)s * MPE_MAX_BLOBS * 2 + i] : 0.0;
    r.centers[i] = (i < 2 * nd) ? a.centers[(size_t)s * MPE_MAX_BLOBS * 2 + i] : 0.f;
  }
  out[s] = r;
}

__global__ void track_reset_kernel(StreamState* st, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  StreamState z;
  for (int i = 0; i < 16; ++i) { const double I = (i % 5 == 0) ? 1.0 : 0.0; z.current_pose[i] = I; z.previous_pose[i] = I; z.predicted_pose[i] = I; }
  z.current_time = 0; z.previous_time = 0; z.predicted_time = 0;
  z.it_since_initialized = 0; z.pad = 0;
  st[s] = z;
}

static inline int grid_for(int n) { return (n + 127) / 128; }

cudaError_t launch_track_begin(const TrackArgs& a, cudaStream_t st) { track_begin_kernel<<<grid_for(a.n), 128, 0, st>>>(a); return cudaGetLastError(); }
cudaError_t launch_track_after_detect(const TrackArgs& a, int pass, cudaStream_t st) {
  if (pass == 0) track_after_detect0_kernel<<<grid_for(a.n), 128, 0, st>>>(a);
  else track_after_detect_kernel<<<grid_for(a.n), 128, 0, st>>>(a, pass);
  return cudaGetLastError();
}
cudaError_t launch_track_finish(const TrackArgs& a, mpe_result* out, cudaStream_t st) { track_finish_kernel<<<grid_for(a.n), 128, 0, st>>>(a, out); return cudaGetLastError(); }
cudaError_t launch_track_reset(StreamState* s, int n, cudaStream_t st) { track_reset_kernel<<<grid_for(n), 128, 0, st>>>(s, n); return cudaGetLastError(); }

}  // namespace mpe
