// K3 — PoseEstimator::checkCorrespondences (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:394-542,
// with calculateSquaredReprojectionErrorAndCertainty :303-342 and the Kabsch fit computeTransformation :908-930)
// followed by PoseEstimator::optimisePose (:733-792: Gauss-Newton on SE(3) with computeJacobian :932-960,
// exponentialMap :962-994, A.ldlt().solve(b), covariance = A^-1 of the last iteration).
//
// Mapping: one WARP per frame.
//   check   lane <-> one 3-subset of the correspondence rows (P3P + greedy matching of the unused rows);
//           contributions to the mean re-projected object points are summed in subset order (lane order),
//           so the floating-point result does not depend on the warp schedule;
//   GN      lane <-> one correspondence: residual, 2x6 Jacobian, J^T J (21 unique entries) and J^T e are
//           formed per lane and accumulated in correspondence order through shuffles into registers that
//           every lane holds; the 6x6 pivoted LDL^T solve and the exponential map are then evaluated
//           redundantly by all lanes (no broadcast, no divergence).
// Latency bound (a dependent chain of ~5 iterations); reads < 1 KB per frame.  Compiled with -fmad=false.
#include "mpe_internal.cuh"
#include "p3p_device.cuh"
#include <cstdio>

namespace mpe {

constexpr int kK3WarpsPerCta = 4;
constexpr int kMaxUnused = MPE_MAX_LEDS - 3;

__device__ __forceinline__ void unrank_comb3_k3(int n, int idx, int& a, int& b, int& c) {
  a = 0;
  for (;;) {
    int m = n - 1 - a;
    int cnt = m * (m - 1) / 2;
    if (idx < cnt) break;
    idx -= cnt;
    ++a;
  }
  b = a + 1;
  for (;;) {
    int cnt = n - 1 - b;
    if (idx < cnt) break;
    idx -= cnt;
    ++b;
  }
  c = b + 1 + idx;
}

// pose_estimator.cpp:303-342.  d is (ni x no) row-major: i image points (rows), j back-projected object
// points (cols).  minCoeff visits column-major and keeps the first strict minimum.
__device__ double squared_error_and_certainty(double* d, int ni, int no, double tol, double* certainty) {
  double squared_error = 0;
  int num = 0;
  int lim = ni < no ? ni : no;
  for (int it = 1; it <= lim; ++it) {
    double mv = d[0];
    int ri = 0, ci = 0;
    for (int j = 0; j < no; ++j)
      for (int i = 0; i < ni; ++i) {
        double v = d[i * no + j];
        if (v < mv) { mv = v; ri = i; ci = j; }
      }
    if (mv <= tol) {
      double v = d[ri * no + ci];
      squared_error += v * v;
      ++num;
      for (int j = 0; j < no; ++j) d[ri * no + j] = HUGE_VAL;
      for (int i = 0; i < ni; ++i) d[i * no + ci] = HUGE_VAL;
    } else {
      break;
    }
  }
  *certainty = (double)num / (double)no;
  return squared_error;
}

// 3x3 one-sided Jacobi SVD, identical operation order to oracle/pose_oracle.cpp svd3 (only + - * / sqrt).
__device__ void svd3(const double Ain[3][3], double U[3][3], double V[3][3]) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { a[i][j] = Ain[i][j]; v[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) { alpha += a[i][p] * a[i][p]; beta += a[i][q] * a[i][q]; gamma += a[i][p] * a[i][q]; }
        if (gamma == 0) continue;
        off = fmax(off, fabs(gamma) / sqrt(alpha * beta));
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; ++i) {
          double ap = a[i][p], aq = a[i][q];
          a[i][p] = c * ap - s * aq; a[i][q] = s * ap + c * aq;
          double vp = v[i][p], vq = v[i][q];
          v[i][p] = c * vp - s * vq; v[i][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-16) break;
  }
  double sv[3];
  for (int j = 0; j < 3; ++j) sv[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
  int order[3] = {0, 1, 2};   // stable descending sort of three values (same result as std::sort on distinct keys)
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (sv[order[j]] < sv[order[j + 1]]) { int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t; }
  for (int jj = 0; jj < 3; ++jj) {
    int j = order[jj];
    for (int i = 0; i < 3; ++i) { V[i][jj] = v[i][j]; U[i][jj] = (sv[j] > 0) ? a[i][j] / sv[j] : 0.0; }
  }
  if (!(sv[order[2]] > 1e-300 * sv[order[0]])) {
    double u0[3] = {U[0][0], U[1][0], U[2][0]}, u1[3] = {U[0][1], U[1][1], U[2][1]};
    U[0][2] = u0[1] * u1[2] - u0[2] * u1[1];
    U[1][2] = u0[2] * u1[0] - u0[0] * u1[2];
    U[2][2] = u0[0] * u1[1] - u0[1] * u1[0];
  }
}

// Symmetric solve by LDL^T with diagonal pivoting; same arithmetic, in the same order, as oracle ldlt_solve6.
// Written with compile-time indices only (the pivot exchange is a predicated swap over the unrolled candidates), so
// the 6x6 system stays in registers.  NB: the first version indexed local arrays with the run-time pivot; nvcc 12.9
// (-O3, sm_100a) miscompiled it inside this kernel (wrong solution, fixed by -Xcicc -O1) — see DESIGN.md, "toolchain".
__device__ __forceinline__ void swap_d(double& a, double& b) { double t = a; a = b; b = t; }

__device__ __forceinline__ void ldlt_solve6(const double Ain[36], const double bin[6], double x[6]) {
  double A[6][6], y[6];
  int piv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    y[i] = bin[i];
#pragma unroll
    for (int j = 0; j < 6; ++j) A[i][j] = Ain[i * 6 + j];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      double v = fabs(A[i][i]);
      if (v > best) { best = v; p = i; }
    }
    piv[k] = p;
#pragma unroll
    for (int q = k + 1; q < 6; ++q) {
      if (p == q) {   // symmetric exchange of rows/columns k and q, right-hand side follows
#pragma unroll
        for (int j = 0; j < 6; ++j) swap_d(A[k][j], A[q][j]);
#pragma unroll
        for (int i = 0; i < 6; ++i) swap_d(A[i][k], A[i][q]);
        swap_d(y[k], y[q]);
      }
    }
    double d = A[k][k];
#pragma unroll
    for (int i = k + 1; i < 6; ++i)
#pragma unroll
      for (int j = k + 1; j <= i; ++j) A[i][j] -= A[i][k] * A[j][k] / d;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
#pragma unroll
    for (int i = k + 1; i < 6; ++i)
#pragma unroll
      for (int j = k + 1; j < i; ++j) A[j][i] = A[i][j];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] /= A[i][i];
#pragma unroll
  for (int i = 5; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
  // undo the exchanges (x = P^T y): transpositions in reverse order
#pragma unroll
  for (int k = 5; k >= 0; --k) {
#pragma unroll
    for (int q = k + 1; q < 6; ++q)
      if (piv[k] == q) swap_d(y[k], y[q]);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// General 6x6 inverse (Gauss-Jordan, partial pivoting); same arithmetic as oracle inverse6, compile-time indices.
__device__ __forceinline__ void inverse6(const double Ain[36], double out[36]) {
  double a[6][12];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) { a[i][j] = Ain[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      double v = fabs(a[i][k]);
      if (v > best) { best = v; p = i; }
    }
#pragma unroll
    for (int q = k + 1; q < 6; ++q) {
      if (p == q) {
#pragma unroll
        for (int j = 0; j < 12; ++j) swap_d(a[k][j], a[q][j]);
      }
    }
    double pv = a[k][k];
#pragma unroll
    for (int j = 0; j < 12; ++j) a[k][j] /= pv;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i != k) {
        double fct = a[i][k];
        if (fct != 0) {
#pragma unroll
          for (int j = 0; j < 12; ++j) a[i][j] -= fct * a[k][j];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) out[i * 6 + j] = a[i][6 + j];
}

// pose_estimator.cpp:962-994; T (3x4 row-major top rows) <- exp(twist) * T
__device__ void exp_map_left_multiply(const double twist[6], double T[12]) {
  double ux = twist[0], uy = twist[1], uz = twist[2];
  double wx = twist[3], wy = twist[4], wz = twist[5];
  double theta = sqrt(wx * wx + wy * wy + wz * wz);
  double theta_squared = theta * theta;
  double O[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
  double O2[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  double rot[3][3], V[3][3];
  if (theta == 0) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    double s = sin(theta), c = cos(theta);
    double kv1 = (1 - c) / (theta_squared), kv2 = (theta - s) / (theta_squared * theta);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double I = (i == j) ? 1.0 : 0.0;
        rot[i][j] = I + O[i][j] / theta * s + O2[i][j] / theta_squared * (1 - c);
        V[i][j] = I + kv1 * O[i][j] + kv2 * O2[i][j];
      }
  }
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = V[i][0] * ux + V[i][1] * uy + V[i][2] * uz;
  // E = [rot t; 0 0 0 1];  T_new = E * [T; 0 0 0 1] with the 4-term dot products of a 4x4 product
  double Tn[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double bottom = (j == 3) ? 1.0 : 0.0;
      double sacc = rot[i][0] * T[j];
      sacc += rot[i][1] * T[4 + j];
      sacc += rot[i][2] * T[8 + j];
      sacc += t[i] * bottom;
      Tn[4 * i + j] = sacc;
    }
  for (int e = 0; e < 12; ++e) T[e] = Tn[e];
}

struct K3Warp {
  double det[MPE_MAX_DET][2];
  double bearing[MPE_MAX_DET][3];
  uint32_t corr[MPE_MAX_LEDS][2];
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(32 * kK3WarpsPerCta) validate_refine_kernel(const K3Args a) {
  __shared__ K3Warp shw[kK3WarpsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kK3WarpsPerCta + warp;
  if (f >= a.n_frames) return;
  if (a.active && !a.active[f]) return;
  K3Warp& sh = shw[warp];
  const int n_obj = a.pp.n_obj;
  const double* K = a.cam.K;

  const int n_det = a.n_det[f];
  int k = a.n_corr[f];
  bool pose_path = (n_det >= 0 && n_det <= MPE_MAX_DET);   // memory-safety guard; n_det < 4 never reaches here with k >= 4 on the cold path
  if (!pose_path) k = 0;
  if (lane < n_det && pose_path) {
    const double* det = a.det + (size_t)f * a.det_stride * 2;
    double u = det[2 * lane], v = det[2 * lane + 1];
    sh.det[lane][0] = u; sh.det[lane][1] = v;
    double x = (u - K[2]) / K[0], y = (v - K[5]) / K[4], z = 1;       // calculateImageVectors :288-301
    double n = sqrt(x * x + y * y + z * z);
    sh.bearing[lane][0] = x / n; sh.bearing[lane][1] = y / n; sh.bearing[lane][2] = z / n;
  }
  if (lane < k) {
    sh.corr[lane][0] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * lane];
    sh.corr[lane][1] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * lane + 1];
  }
  __syncwarp();

  double T[12];   // predicted_pose_ top three rows, row-major
  for (int e = 0; e < 12; ++e) T[e] = a.pose_io[(size_t)f * 16 + e];
  int ok = 0;

  // ------------------------------------------------------------------ checkCorrespondences
  if (a.mode == 0 || a.mode == 1) {
    if (k >= 4) {                                                      // :401
      const int N = k * (k - 1) * (k - 2) / 6;
      const int nu = k - 3;                                            // total_unused_correspondences
      double mean[MPE_MAX_LEDS][3];
      for (int j = 0; j < n_obj; ++j) mean[j][0] = mean[j][1] = mean[j][2] = 0;
      int num_valid = 0;
      for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        int found = 0;
        double contrib[MPE_MAX_LEDS][3];
        if (i < N) {
          int c0, c1, c2;
          unrank_comb3_k3(k, i, c0, c1, c2);
          const int l0 = sh.corr[c0][0] - 1, l1 = sh.corr[c1][0] - 1, l2 = sh.corr[c2][0] - 1;
          const int e0 = sh.corr[c0][1] - 1, e1 = sh.corr[c1][1] - 1, e2 = sh.corr[c2][1] - 1;
          const double* mk = a.pp.markers;
          P3PSetup S;
          int rc = p3p_setup(v_make(sh.bearing[e0][0], sh.bearing[e0][1], sh.bearing[e0][2]),
                             v_make(sh.bearing[e1][0], sh.bearing[e1][1], sh.bearing[e1][2]),
                             v_make(sh.bearing[e2][0], sh.bearing[e2][1], sh.bearing[e2][2]),
                             v_make(mk[3 * l0], mk[3 * l0 + 1], mk[3 * l0 + 2]), v_make(mk[3 * l1], mk[3 * l1 + 1], mk[3 * l1 + 2]),
                             v_make(mk[3 * l2], mk[3 * l2 + 1], mk[3 * l2 + 2]), S);
          if (rc == 0) {
            double min_sq = HUGE_VAL;
            int best = 0;
            for (int j = 0; j < 4; ++j) {
              double H[12];
              p3p_solution(S, j, H);
              if (!h_is_finite(H)) continue;                           // :479
              double Hi[12], KT[12];
              h_inverse(H, Hi);
              kt_product(K, Hi, KT);
              double bu[kMaxUnused], bv[kMaxUnused];
              double dist[kMaxUnused * kMaxUnused];
              int m = 0;
              for (int l = 0; l < k; ++l) {                            // unused rows, in row order (:437-455)
                if (l == c0 || l == c1 || l == c2) continue;
                int led = sh.corr[l][0] - 1;
                kt_project(KT, mk[3 * led], mk[3 * led + 1], mk[3 * led + 2], bu[m], bv[m]);
                ++m;
              }
              int ii = 0;
              for (int l = 0; l < k; ++l) {
                if (l == c0 || l == c1 || l == c2) continue;
                int di = sh.corr[l][1] - 1;
                for (int jj = 0; jj < nu; ++jj) {
                  double dx = sh.det[di][0] - bu[jj], dy = sh.det[di][1] - bv[jj];
                  dist[ii * nu + jj] = sqrt(dx * dx + dy * dy);
                }
                ++ii;
              }
              double certainty;
              double sq = squared_error_and_certainty(dist, nu, nu, a.pp.back_projection_pixel_tolerance, &certainty);
              if (certainty >= a.pp.certainty_threshold) {             // :494
                found = 1;
                if (sq < min_sq) { min_sq = sq; best = j; }
              }
            }
            if (found) {                                               // :506-518
              double H[12], Hi[12];
              p3p_solution(S, best, H);
              h_inverse(H, Hi);
              for (int jj = 0; jj < n_obj; ++jj) {
                double x = mk[3 * jj], y = mk[3 * jj + 1], z = mk[3 * jj + 2];
                for (int r = 0; r < 3; ++r) {
                  double sacc = Hi[4 * r] * x;
                  sacc += Hi[4 * r + 1] * y;
                  sacc += Hi[4 * r + 2] * z;
                  sacc += Hi[4 * r + 3] * 1.0;
                  contrib[jj][r] = sacc;
                }
              }
            }
          }
        }
        // ordered accumulation over the subsets of this pass
        unsigned vmask = __ballot_sync(0xffffffffu, found != 0);
        num_valid += __popc(vmask);
        while (vmask) {
          int src = __ffs(vmask) - 1;
          vmask &= vmask - 1;
          for (int jj = 0; jj < n_obj; ++jj)
            for (int r = 0; r < 3; ++r) mean[jj][r] = mean[jj][r] + shfl_d(contrib[jj][r], src);
        }
      }
      if ((double)num_valid / N >= a.pp.valid_correspondence_threshold) {   // :525
        ok = 1;
        // computeTransformation (:908-930)
        const double* mk = a.pp.markers;
        double mo[3] = {0, 0, 0}, mr[3] = {0, 0, 0};
        for (int j = 0; j < n_obj; ++j)
          for (int r = 0; r < 3; ++r) {
            mean[j][r] = mean[j][r] / num_valid;
            mo[r] = mo[r] + mk[3 * j + r];
            mr[r] = mr[r] + mean[j][r];
          }
        for (int r = 0; r < 3; ++r) { mo[r] = mo[r] / (double)n_obj; mr[r] = mr[r] / (double)n_obj; }
        double Hm[3][3], U[3][3], V[3][3];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double sacc = 0;
            for (int j = 0; j < n_obj; ++j) sacc += (mk[3 * j + r] - mo[r]) * (mean[j][c] - mr[c]);
            Hm[r][c] = sacc;
          }
        svd3(Hm, U, V);
        double Rm[3][3];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) Rm[r][c] = V[r][0] * U[c][0] + V[r][1] * U[c][1] + V[r][2] * U[c][2];   // V * U^T
        for (int r = 0; r < 3; ++r) {
          double rt = Rm[r][0] * mo[0] + Rm[r][1] * mo[1] + Rm[r][2] * mo[2];
          T[4 * r] = Rm[r][0]; T[4 * r + 1] = Rm[r][1]; T[4 * r + 2] = Rm[r][2];
          T[4 * r + 3] = mr[r] - rt;
        }
      }
    }
  } else {
    ok = 1;   // mode 2: caller supplies correspondences and the starting pose
  }

  // ------------------------------------------------------------------ optimisePose
  int iters = 0;
  double cov[36];
  bool ran_gn = false;
  if (ok && (a.mode == 0 || a.mode == 2)) {
    ran_gn = true;
    const double fx = K[0], fy = K[4];
    const double* mk = a.pp.markers;
    const bool lane_active = (lane < k) && (sh.corr[lane < k ? lane : 0][1] != 0);   // :761
    double ox = 0, oy = 0, oz = 0, ix = 0, iy = 0;
    if (lane_active) {
      int led = sh.corr[lane][0] - 1, di = sh.corr[lane][1] - 1;
      ox = mk[3 * led]; oy = mk[3 * led + 1]; oz = mk[3 * led + 2];
      ix = sh.det[di][0]; iy = sh.det[di][1];
    }
    const unsigned amask = __ballot_sync(0xffffffffu, lane_active);
    double A[36], b[6], dT[6];
    for (int e = 0; e < 36; ++e) A[e] = 0;
    for (int it = 0; it < 500; ++it) {                                  // max_itr :738
      double jt[27];   // 21 unique J^T J entries (upper triangle, row-major) + 6 J^T e entries
      if (lane_active) {
        double KT[12], pu, pv;
        kt_product(K, T, KT);                                           // project2d :251-268
        kt_project(KT, ox, oy, oz, pu, pv);
        double e0 = ix - pu, e1 = iy - pv;                              // :769
        // computeJacobian :932-960
        double x = T[0] * ox + T[1] * oy + T[2] * oz + T[3] * 1.0;
        double y = T[4] * ox + T[5] * oy + T[6] * oz + T[7] * 1.0;
        double z = T[8] * ox + T[9] * oy + T[10] * oz + T[11] * 1.0;
        double z_2 = z * z;
        double J0[6], J1[6];
        J0[0] = 1 / z * fx; J0[1] = 0; J0[2] = -x / z_2 * fx; J0[3] = -x * y / z_2 * fx; J0[4] = (1 + (x * x / z_2)) * fx; J0[5] = -y / z * fx;
        J1[0] = 0; J1[1] = 1 / z * fy; J1[2] = -y / z_2 * fy; J1[3] = -(1 + y * y / z_2) * fy; J1[4] = x * y / z_2 * fy; J1[5] = x / z * fy;
        int q = 0;
        for (int r = 0; r < 6; ++r)
          for (int c = r; c < 6; ++c) jt[q++] = J0[r] * J0[c] + J1[r] * J1[c];
        for (int r = 0; r < 6; ++r) jt[21 + r] = J0[r] * e0 + J1[r] * e1;
      } else {
        for (int q = 0; q < 27; ++q) jt[q] = 0;
      }
      double acc[27];
      for (int q = 0; q < 27; ++q) acc[q] = 0;
      unsigned mm = amask;
      while (mm) {                                                      // correspondence order (:759)
        int src = __ffs(mm) - 1;
        mm &= mm - 1;
        for (int q = 0; q < 27; ++q) acc[q] += shfl_d(jt[q], src);
      }
      {
        int q = 0;
        for (int r = 0; r < 6; ++r)
          for (int c = r; c < 6; ++c) { A[r * 6 + c] = acc[q]; A[c * 6 + r] = acc[q]; ++q; }
        for (int r = 0; r < 6; ++r) b[r] = acc[21 + r];
      }
      ldlt_solve6(A, b, dT);                                            // :778
      exp_map_left_multiply(dT, T);                                     // :781
      ++iters;
      double mx = -1;                                                   // norm_max :1073-1085
      for (int q = 0; q < 6; ++q) { double av = fabs(dT[q]); if (av > mx) mx = av; }
#ifdef MPE_DEBUG_GN
      if (lane == 0 && f == 0 && iters == 1) {
        for (int r = 0; r < 6; ++r) printf("A[%d] %.9e %.9e %.9e %.9e %.9e %.9e | b %.9e\n", r, A[r*6], A[r*6+1], A[r*6+2], A[r*6+3], A[r*6+4], A[r*6+5], b[r]);
        printf("amask %x k %d\n", amask, k);
      }
      if (lane == 0 && f == 0) printf("GN it %d mx %.3e dT %.3e %.3e %.3e %.3e %.3e %.3e | A00 %.6e A55 %.6e b0 %.3e T3 %.9f %.9f %.9f\n", iters, mx, dT[0], dT[1], dT[2], dT[3], dT[4], dT[5], A[0], A[35], b[0], T[3], T[7], T[11]);
#endif
      if (mx <= 1e-13) break;                                           // :786
    }
    inverse6(A, cov);                                                   // :790
  }

  if (lane == 0) {
    double* po = a.pose_io + (size_t)f * 16;
    if (ok) {
      for (int e = 0; e < 12; ++e) po[e] = T[e];
      po[12] = 0; po[13] = 0; po[14] = 0; po[15] = 1;
    }
    if (ran_gn && a.cov) for (int e = 0; e < 36; ++e) a.cov[(size_t)f * 36 + e] = cov[e];
    if (a.ok) a.ok[f] = ok;
    if (a.iters) a.iters[f] = iters;
    if (a.updated) a.updated[f] = (ok && ran_gn) ? 1 : 0;
  }
}

cudaError_t launch_validate_refine(const K3Args& a, cudaStream_t st) {
  int grid = (a.n_frames + kK3WarpsPerCta - 1) / kK3WarpsPerCta;
  validate_refine_kernel<<<grid, 32 * kK3WarpsPerCta, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace mpe
