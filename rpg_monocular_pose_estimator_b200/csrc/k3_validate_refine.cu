// K3 — PoseEstimator::checkCorrespondences (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:394-542,
// with calculateSquaredReprojectionErrorAndCertainty :303-342 and the Kabsch fit computeTransformation :908-930)
// followed by PoseEstimator::optimisePose (:733-792: Gauss-Newton on SE(3) with computeJacobian :932-960,
// exponentialMap :962-994, A.ldlt().solve(b), covariance = A^-1 of the last iteration).
//
// Mapping (throughput first: thousands of frames per launch, so the parallel axis is the FRAME, not the lane):
//   check_kernel   one thread per (frame, 3-subset of the correspondence rows): P3P + greedy matching of the unused
//                  rows; the subsets' contributions to the mean re-projected object points go through shared memory
//                  and are summed by one thread per frame in subset order (= the reference's loop order), so the
//                  floating-point result is schedule independent;
//   refine_kernel  one thread per frame: acceptance test, Kabsch (3x3 Jacobi SVD), then Gauss-Newton with every
//                  accumulation in correspondence order, pivoted LDL^T, exponential map, covariance = A^-1.
// (The first version used one warp per frame with lanes = subsets / correspondences and shuffle reductions; 31/32 of
//  the issue slots of the sequential parts were wasted: 150 ns/frame against ~15 for this layout.)
// Small batches (<= kGnCooperativeMaxFrames frames, e.g. one camera) are the opposite regime: the duration of a kernel is one
// thread's dependent chain, so they get lane-cooperative kernels with the same arithmetic per value:
//   check_wide_kernel    CTA per frame, warp = P3P solution, lane = subset;
//   gauss_newton_kernel  32 lanes per frame: lane = correspondence for the Jacobians, lane = entry of the normal equations,
//                        lane-parallel pivoted LDL^T and Gauss-Jordan (one frame: 105 -> 37 us).
// FP64 / latency bound; reads < 1 KB per frame.  Compiled with -fmad=false.
#include "mpe_internal.cuh"
#include "p3p_device.cuh"
#include <cstdio>
#include <cstdlib>

namespace mpe {

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

// Symmetric solve by LDL^T with diagonal pivoting; same arithmetic, in the same order, as oracle ldlt_solve6 (which is the
// scheme Eigen's LDLT documents and oracle/eigen_shim implements: lower triangle only, left-looking, the pivot of step k
// is the largest |diagonal| among the not yet updated entries k..5; solve = P, L, D with a zero-pivot guard, L^T, P^T).
// Written with compile-time indices only (the pivot exchange is a predicated swap over the unrolled candidates), so
// the 6x6 system stays in registers.  NB: the first version indexed local arrays with the run-time pivot; nvcc 12.9
// (-O3, sm_100a) miscompiled it inside this kernel (wrong solution, fixed by -Xcicc -O1) — see DESIGN.md, "toolchain".
__device__ __forceinline__ void swap_d(double& a, double& b) { double t = a; a = b; b = t; }

__device__ __forceinline__ void ldlt_solve6(const double Ain[36], const double bin[6], double x[6]) {
  double m[6][6], y[6];   // only m[i][j], i >= j, is used
  int tr[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    y[i] = bin[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) m[i][j] = Ain[i * 6 + j];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int big = k;
    double best = fabs(m[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const double v = fabs(m[i][i]);
      if (v > best) { best = v; big = i; }
    }
    tr[k] = big;
#pragma unroll
    for (int q = k + 1; q < 6; ++q) {
      if (big == q) {   // symmetric exchange of rows/columns k and q inside the lower triangle; the right-hand side follows
#pragma unroll
        for (int j = 0; j < k; ++j) swap_d(m[k][j], m[q][j]);
#pragma unroll
        for (int i = q + 1; i < 6; ++i) swap_d(m[i][k], m[i][q]);
        swap_d(m[k][k], m[q][q]);
#pragma unroll
        for (int i = k + 1; i < q; ++i) swap_d(m[i][k], m[q][i]);
        swap_d(y[k], y[q]);
      }
    }
    if (k > 0) {
      double temp[6];
#pragma unroll
      for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
      double s = m[k][0] * temp[0];
#pragma unroll
      for (int j = 1; j < k; ++j) s += m[k][j] * temp[j];
      m[k][k] -= s;
#pragma unroll
      for (int i = k + 1; i < 6; ++i) {
        double t = m[i][0] * temp[0];
#pragma unroll
        for (int j = 1; j < k; ++j) t += m[i][j] * temp[j];
        m[i][k] -= t;
      }
    }
    const double akk = m[k][k];
    if (fabs(akk) > 0.0) {
#pragma unroll
      for (int i = k + 1; i < 6; ++i) m[i][k] /= akk;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) y[i] -= m[i][j] * y[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) { if (fabs(m[i][i]) > 2.2250738585072014e-308) y[i] /= m[i][i]; else y[i] = 0.0; }
#pragma unroll
  for (int i = 5; i >= 0; --i)
#pragma unroll
    for (int j = 5; j > i; --j) y[i] -= m[j][i] * y[j];
  // undo the exchanges (x = P^T y): transpositions in reverse order
#pragma unroll
  for (int k = 5; k >= 0; --k) {
#pragma unroll
    for (int q = k + 1; q < 6; ++q)
      if (tr[k] == q) swap_d(y[k], y[q]);
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

// ------------------------------------------------------------------------------------------------
// K3a — check_kernel: one THREAD per (frame, 3-subset of the correspondence rows).
// ------------------------------------------------------------------------------------------------
constexpr int kK3aThreads = 128;

struct CheckFrame {
  double det[MPE_MAX_DET][2];
  double bearing[MPE_MAX_DET][3];
  uint32_t corr[MPE_MAX_LEDS][2];
  int k, N, valid;
};

__global__ void __launch_bounds__(kK3aThreads) check_kernel(const K3Args a, int G, int Nmax) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t k3_smem[];
  const int n_obj = a.pp.n_obj;
  CheckFrame* fr = reinterpret_cast<CheckFrame*>(k3_smem);
  double* contrib = reinterpret_cast<double*>(k3_smem + (size_t)G * sizeof(CheckFrame));   // [kK3aThreads][n_obj*3]
  int* found_s = reinterpret_cast<int*>(contrib + (size_t)kK3aThreads * n_obj * 3);       // [kK3aThreads]
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * G;
  const double* K = a.cam.K;
  const double* mk = a.pp.markers;

  // ---- a CTA none of whose frames takes part leaves at once (masked passes of the tracking step: ~0.5 % of the streams re-initialise)
  {
    const int fq = f0 + tid;
    const int takes_part = (tid < G) && (fq < a.n_frames) && !(a.active && !a.active[fq]);
    if (!__syncthreads_or(takes_part)) return;
  }
  // ---- stage the frames of this CTA
  if (tid < G) {
    int f = f0 + tid;
    int valid = (f < a.n_frames) && !(a.active && !a.active[f]);
    int k = 0, n_det = 0;
    if (valid) {
      n_det = a.n_det[f];
      k = a.n_corr[f];
      if (n_det < 0 || n_det > MPE_MAX_DET) k = 0;      // memory-safety guard
      if (k > MPE_MAX_LEDS) k = MPE_MAX_LEDS;
    }
    fr[tid].valid = valid;
    fr[tid].k = k;
    fr[tid].N = (k >= 4) ? k * (k - 1) * (k - 2) / 6 : 0;                  // pose_estimator.cpp:401
  }
  for (int idx = tid; idx < G * MPE_MAX_DET; idx += kK3aThreads) {
    int g = idx / MPE_MAX_DET, l = idx - g * MPE_MAX_DET;
    int f = f0 + g;
    if (f < a.n_frames) {
      int n_det = a.n_det[f];
      if (l < n_det && n_det <= MPE_MAX_DET) {
        const double* det = a.det + (size_t)f * a.det_stride * 2;
        double u = det[2 * l], v = det[2 * l + 1];
        fr[g].det[l][0] = u; fr[g].det[l][1] = v;
        double x = (u - K[2]) / K[0], y = (v - K[5]) / K[4], z = 1;       // calculateImageVectors :288-301
        double n = sqrt(x * x + y * y + z * z);
        fr[g].bearing[l][0] = x / n; fr[g].bearing[l][1] = y / n; fr[g].bearing[l][2] = z / n;
      }
      int k = a.n_corr[f];
      if (l < k && l < MPE_MAX_LEDS) {
        fr[g].corr[l][0] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * l];
        fr[g].corr[l][1] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * l + 1];
      }
    }
  }
  __syncthreads();

  const int g = (G > 1) ? tid / Nmax : 0;
  const bool in_group = (g < G);
  const bool leader = in_group && ((G > 1) ? (tid - g * Nmax == 0) : (tid == 0));
  const int n_pass = (G > 1) ? 1 : (Nmax + kK3aThreads - 1) / kK3aThreads;
  double mean[MPE_MAX_LEDS][3];
  int num_valid = 0;
  if (leader)
    for (int j = 0; j < n_obj; ++j) mean[j][0] = mean[j][1] = mean[j][2] = 0;

  for (int pass = 0; pass < n_pass; ++pass) {
    const int i = (G > 1) ? (tid - g * Nmax) : (pass * kK3aThreads + tid);
    int found = 0;
    if (in_group && fr[g].valid && i < fr[g].N) {
      const CheckFrame& F = fr[g];
      const int k = F.k, nu = k - 3;
      int c0, c1, c2;
      unrank_comb3_k3(k, i, c0, c1, c2);
      const int l0 = F.corr[c0][0] - 1, l1 = F.corr[c1][0] - 1, l2 = F.corr[c2][0] - 1;
      const int e0 = F.corr[c0][1] - 1, e1 = F.corr[c1][1] - 1, e2 = F.corr[c2][1] - 1;
      P3PSetup S;
      int rc = p3p_setup(v_make(F.bearing[e0][0], F.bearing[e0][1], F.bearing[e0][2]),
                         v_make(F.bearing[e1][0], F.bearing[e1][1], F.bearing[e1][2]),
                         v_make(F.bearing[e2][0], F.bearing[e2][1], F.bearing[e2][2]),
                         v_make(mk[3 * l0], mk[3 * l0 + 1], mk[3 * l0 + 2]), v_make(mk[3 * l1], mk[3 * l1 + 1], mk[3 * l1 + 2]),
                         v_make(mk[3 * l2], mk[3 * l2 + 1], mk[3 * l2 + 2]), S);
      if (rc == 0) {
        double min_sq = HUGE_VAL;
        int best = 0;
        for (int j = 0; j < 4; ++j) {
          double H[12];
          if (!p3p_solution(S, j, H)) continue;
          if (!h_is_finite(H)) continue;                           // :479
          double Hi[12], KT[12];
          h_inverse(H, Hi);
          kt_product(K, Hi, KT);
          double bu[kMaxUnused], bv[kMaxUnused];
          double dist[kMaxUnused * kMaxUnused];
          int m = 0;
          for (int l = 0; l < k; ++l) {                            // unused rows, in row order (:437-455)
            if (l == c0 || l == c1 || l == c2) continue;
            int led = F.corr[l][0] - 1;
            kt_project(KT, mk[3 * led], mk[3 * led + 1], mk[3 * led + 2], bu[m], bv[m]);
            ++m;
          }
          int ii = 0;
          for (int l = 0; l < k; ++l) {
            if (l == c0 || l == c1 || l == c2) continue;
            int di = F.corr[l][1] - 1;
            for (int jj = 0; jj < nu; ++jj) {
              double dx = F.det[di][0] - bu[jj], dy = F.det[di][1] - bv[jj];
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
          double* out = contrib + (size_t)tid * n_obj * 3;
          for (int jj = 0; jj < n_obj; ++jj) {
            double x = mk[3 * jj], y = mk[3 * jj + 1], z = mk[3 * jj + 2];
            for (int r = 0; r < 3; ++r) {
              double sacc = Hi[4 * r] * x;
              sacc += Hi[4 * r + 1] * y;
              sacc += Hi[4 * r + 2] * z;
              sacc += Hi[4 * r + 3] * 1.0;
              out[jj * 3 + r] = sacc;
            }
          }
        }
      }
    }
    found_s[tid] = found;
    __syncthreads();
    if (leader && fr[g].valid) {
      // ordered accumulation over the subsets of this pass (subset order = reference loop order)
      const int base = (G > 1) ? g * Nmax : 0;
      const int cnt = (G > 1) ? fr[g].N : min(kK3aThreads, fr[g].N - pass * kK3aThreads);
      for (int j = 0; j < cnt; ++j) {
        if (!found_s[base + j]) continue;
        ++num_valid;
        const double* cj = contrib + (size_t)(base + j) * n_obj * 3;
        for (int jj = 0; jj < n_obj; ++jj)
          for (int r = 0; r < 3; ++r) mean[jj][r] = mean[jj][r] + cj[jj * 3 + r];
      }
    }
    __syncthreads();
  }
  if (leader && fr[g].valid) {
    const int f = f0 + g;
    double* so = a.check_sums + (size_t)f * MPE_MAX_LEDS * 3;
    for (int jj = 0; jj < n_obj; ++jj)
      for (int r = 0; r < 3; ++r) so[jj * 3 + r] = mean[jj][r];
    a.check_cnt[2 * f] = num_valid;
    a.check_cnt[2 * f + 1] = fr[g].N;
  }
}

// ------------------------------------------------------------------------------------------------
// K3a' — check_wide_kernel: the same computation for SMALL batches (single cameras).  With a thread per subset the
// kernel's duration is one thread's chain: P3P set-up + four back-substitutions and scorings (61 us for one frame).  Here a
// CTA serves one frame, warp j evaluates solution j of every subset (lane = subset), so the chain is set-up + ONE solution;
// warp 0 then picks each subset's best valid solution in solution order (same strict '<' rule), forms its contribution, and
// one thread adds the contributions in subset order.  Identical arithmetic per value, hence identical results.
// ------------------------------------------------------------------------------------------------
constexpr int kWideThreads = 128;     // 4 warps = the 4 P3P solutions

__global__ void __launch_bounds__(kWideThreads) check_wide_kernel(const K3Args a) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t k3w_smem[];
  const int n_obj = a.pp.n_obj;
  CheckFrame& F = *reinterpret_cast<CheckFrame*>(k3w_smem);
  double* sq_s = reinterpret_cast<double*>(k3w_smem + sizeof(CheckFrame));          // [4][32] squared error of solution j, subset lane
  int* valid_s = reinterpret_cast<int*>(sq_s + 4 * 32);                              // [4][32] certainty >= threshold
  double* contrib = reinterpret_cast<double*>(valid_s + 4 * 32);                     // [32][n_obj*3]
  int* found_s = reinterpret_cast<int*>(contrib + (size_t)32 * n_obj * 3);           // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.x;
  const double* K = a.cam.K;
  const double* mk = a.pp.markers;
  if (a.active && !a.active[f]) return;                       // uniform over the CTA

  if (tid == 0) {
    int n_det = a.n_det[f], k = a.n_corr[f];
    if (n_det < 0 || n_det > MPE_MAX_DET) k = 0;
    if (k > MPE_MAX_LEDS) k = MPE_MAX_LEDS;
    F.valid = 1; F.k = k;
    F.N = (k >= 4) ? k * (k - 1) * (k - 2) / 6 : 0;           // pose_estimator.cpp:401
  }
  if (tid < MPE_MAX_DET) {
    const int l = tid;
    const int n_det = a.n_det[f];
    if (l < n_det && n_det <= MPE_MAX_DET) {
      const double* det = a.det + (size_t)f * a.det_stride * 2;
      double u = det[2 * l], v = det[2 * l + 1];
      F.det[l][0] = u; F.det[l][1] = v;
      double x = (u - K[2]) / K[0], y = (v - K[5]) / K[4], z = 1;         // calculateImageVectors :288-301
      double n = sqrt(x * x + y * y + z * z);
      F.bearing[l][0] = x / n; F.bearing[l][1] = y / n; F.bearing[l][2] = z / n;
    }
    const int k = a.n_corr[f];
    if (l < k && l < MPE_MAX_LEDS) {
      F.corr[l][0] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * l];
      F.corr[l][1] = a.corr[(size_t)f * 2 * MPE_MAX_LEDS + 2 * l + 1];
    }
  }
  __syncthreads();

  const int N = F.N, k = F.k, nu = k - 3;
  const int n_pass = (N + 31) / 32;
  double mean[MPE_MAX_LEDS][3];
  int num_valid = 0;
  if (tid == 0)
    for (int j = 0; j < n_obj; ++j) mean[j][0] = mean[j][1] = mean[j][2] = 0;

  for (int pass = 0; pass < n_pass; ++pass) {
    const int i = pass * 32 + lane;                            // subset
    const int j = warp;                                        // solution
    P3PSetup S;
    int rc = -1, c0 = 0, c1 = 0, c2 = 0;
    double sq = 0;
    int valid = 0;
    if (i < N) {
      unrank_comb3_k3(k, i, c0, c1, c2);
      const int l0 = F.corr[c0][0] - 1, l1 = F.corr[c1][0] - 1, l2 = F.corr[c2][0] - 1;
      const int e0 = F.corr[c0][1] - 1, e1 = F.corr[c1][1] - 1, e2 = F.corr[c2][1] - 1;
      rc = p3p_setup(v_make(F.bearing[e0][0], F.bearing[e0][1], F.bearing[e0][2]),
                     v_make(F.bearing[e1][0], F.bearing[e1][1], F.bearing[e1][2]),
                     v_make(F.bearing[e2][0], F.bearing[e2][1], F.bearing[e2][2]),
                     v_make(mk[3 * l0], mk[3 * l0 + 1], mk[3 * l0 + 2]), v_make(mk[3 * l1], mk[3 * l1 + 1], mk[3 * l1 + 2]),
                     v_make(mk[3 * l2], mk[3 * l2 + 1], mk[3 * l2 + 2]), S);
      if (rc == 0) {
        double H[12];
        if (p3p_solution(S, j, H) && h_is_finite(H)) {           // :479
          double Hi[12], KT[12];
          h_inverse(H, Hi);
          kt_product(K, Hi, KT);
          double bu[kMaxUnused], bv[kMaxUnused];
          double dist[kMaxUnused * kMaxUnused];
          int m = 0;
          for (int l = 0; l < k; ++l) {                          // unused rows, in row order (:437-455)
            if (l == c0 || l == c1 || l == c2) continue;
            int led = F.corr[l][0] - 1;
            kt_project(KT, mk[3 * led], mk[3 * led + 1], mk[3 * led + 2], bu[m], bv[m]);
            ++m;
          }
          int ii = 0;
          for (int l = 0; l < k; ++l) {
            if (l == c0 || l == c1 || l == c2) continue;
            int di = F.corr[l][1] - 1;
            for (int jj = 0; jj < nu; ++jj) {
              double dx = F.det[di][0] - bu[jj], dy = F.det[di][1] - bv[jj];
              dist[ii * nu + jj] = sqrt(dx * dx + dy * dy);
            }
            ++ii;
          }
          double certainty;
          sq = squared_error_and_certainty(dist, nu, nu, a.pp.back_projection_pixel_tolerance, &certainty);
          valid = (certainty >= a.pp.certainty_threshold) ? 1 : 0;  // :494
        }
      }
    }
    sq_s[j * 32 + lane] = sq;
    valid_s[j * 32 + lane] = valid;
    __syncthreads();
    if (warp == 0) {
      int found = 0;
      if (i < N && rc == 0) {
        double min_sq = HUGE_VAL;
        int best = 0;
        for (int jj = 0; jj < 4; ++jj) {
          if (!valid_s[jj * 32 + lane]) continue;
          found = 1;
          const double sj = sq_s[jj * 32 + lane];
          if (sj < min_sq) { min_sq = sj; best = jj; }
        }
        if (found) {                                             // :506-518
          double H[12], Hi[12];
          p3p_solution(S, best, H);
          h_inverse(H, Hi);
          double* out = contrib + (size_t)lane * n_obj * 3;
          for (int jj = 0; jj < n_obj; ++jj) {
            double x = mk[3 * jj], y = mk[3 * jj + 1], z = mk[3 * jj + 2];
            for (int r = 0; r < 3; ++r) {
              double sacc = Hi[4 * r] * x;
              sacc += Hi[4 * r + 1] * y;
              sacc += Hi[4 * r + 2] * z;
              sacc += Hi[4 * r + 3] * 1.0;
              out[jj * 3 + r] = sacc;
            }
          }
        }
      }
      found_s[lane] = found;
    }
    __syncthreads();
    if (tid == 0) {
      const int cnt = min(32, N - pass * 32);
      for (int jx = 0; jx < cnt; ++jx) {
        if (!found_s[jx]) continue;
        ++num_valid;
        const double* cj = contrib + (size_t)jx * n_obj * 3;
        for (int jj = 0; jj < n_obj; ++jj)
          for (int r = 0; r < 3; ++r) mean[jj][r] = mean[jj][r] + cj[jj * 3 + r];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    double* so = a.check_sums + (size_t)f * MPE_MAX_LEDS * 3;
    for (int jj = 0; jj < n_obj; ++jj)
      for (int r = 0; r < 3; ++r) so[jj * 3 + r] = mean[jj][r];
    a.check_cnt[2 * f] = num_valid;
    a.check_cnt[2 * f + 1] = N;
  }
}

// ------------------------------------------------------------------------------------------------
// K3b — refine_kernel: one THREAD per frame: acceptance test + Kabsch of checkCorrespondences, then the
// Gauss-Newton of optimisePose, every sum in the reference's loop order.
// ------------------------------------------------------------------------------------------------
constexpr int kK3bThreads = 64;

__global__ void __launch_bounds__(kK3bThreads) refine_kernel(const K3Args a) {
  pdl_enter();
  const int f = blockIdx.x * kK3bThreads + threadIdx.x;
  if (f >= a.n_frames) return;
  if (a.active && !a.active[f]) return;
  const int n_obj = a.pp.n_obj;
  const double* K = a.cam.K;
  const double* mk = a.pp.markers;
  const int n_det = a.n_det[f];
  int k = a.n_corr[f];
  if (n_det < 0 || n_det > MPE_MAX_DET) k = 0;
  if (k > MPE_MAX_LEDS) k = MPE_MAX_LEDS;
  const double* det = a.det + (size_t)f * a.det_stride * 2;
  const uint32_t* corr = a.corr + (size_t)f * 2 * MPE_MAX_LEDS;

  double T[12];   // predicted_pose_ top three rows, row-major
  for (int e = 0; e < 12; ++e) T[e] = a.pose_io[(size_t)f * 16 + e];
  int ok = 0;

  if (a.mode == 0 || a.mode == 1) {
    const int num_valid = a.check_cnt[2 * f], N = a.check_cnt[2 * f + 1];
    if (N > 0 && (double)num_valid / N >= a.pp.valid_correspondence_threshold) {   // :525
      ok = 1;
      const double* sums = a.check_sums + (size_t)f * MPE_MAX_LEDS * 3;
      // computeTransformation (:908-930)
      double mo[3] = {0, 0, 0}, mr[3] = {0, 0, 0};
      for (int j = 0; j < n_obj; ++j)
        for (int r = 0; r < 3; ++r) {
          mo[r] = mo[r] + mk[3 * j + r];
          mr[r] = mr[r] + sums[j * 3 + r] / num_valid;
        }
      for (int r = 0; r < 3; ++r) { mo[r] = mo[r] / (double)n_obj; mr[r] = mr[r] / (double)n_obj; }
      double Hm[3][3], U[3][3], V[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double sacc = 0;
          for (int j = 0; j < n_obj; ++j) sacc += (mk[3 * j + r] - mo[r]) * (sums[j * 3 + c] / num_valid - mr[c]);
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
  } else {
    ok = 1;   // mode 2: caller supplies correspondences and the starting pose
  }

  int iters = 0;
  bool ran_gn = false;
  double A[36];
  if (ok && (a.mode == 0 || a.mode == 2)) {
    ran_gn = true;
    const double fx = K[0], fy = K[4];
    double b[6], dT[6];
    for (int e = 0; e < 36; ++e) A[e] = 0;
    for (int it = 0; it < 500; ++it) {                                  // max_itr :738
      for (int e = 0; e < 36; ++e) A[e] = 0;
      for (int e = 0; e < 6; ++e) b[e] = 0;
      double KT[12];
      kt_product(K, T, KT);                                             // project2d :251-268 (same KT for every point)
      for (int j = 0; j < k; ++j) {                                     // :759
        const uint32_t led1 = corr[2 * j], det1 = corr[2 * j + 1];
        if (det1 == 0) continue;                                        // :761
        const int led = (int)led1 - 1, di = (int)det1 - 1;
        const double ox = mk[3 * led], oy = mk[3 * led + 1], oz = mk[3 * led + 2];
        double pu, pv;
        kt_project(KT, ox, oy, oz, pu, pv);
        const double e0 = det[2 * di] - pu, e1 = det[2 * di + 1] - pv;  // :769
        // computeJacobian :932-960
        double x = T[0] * ox + T[1] * oy + T[2] * oz + T[3] * 1.0;
        double y = T[4] * ox + T[5] * oy + T[6] * oz + T[7] * 1.0;
        double z = T[8] * ox + T[9] * oy + T[10] * oz + T[11] * 1.0;
        double z_2 = z * z;
        double J0[6], J1[6];
        J0[0] = 1 / z * fx; J0[1] = 0; J0[2] = -x / z_2 * fx; J0[3] = -x * y / z_2 * fx; J0[4] = (1 + (x * x / z_2)) * fx; J0[5] = -y / z * fx;
        J1[0] = 0; J1[1] = 1 / z * fy; J1[2] = -y / z_2 * fy; J1[3] = -(1 + y * y / z_2) * fy; J1[4] = x * y / z_2 * fy; J1[5] = x / z * fy;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int c = 0; c < 6; ++c) A[r * 6 + c] += J0[r] * J0[c] + J1[r] * J1[c];
          b[r] += J0[r] * e0 + J1[r] * e1;
        }
      }
      ldlt_solve6(A, b, dT);                                            // :778
      exp_map_left_multiply(dT, T);                                     // :781
      ++iters;
      double mx = -1;                                                   // norm_max :1073-1085
#pragma unroll
      for (int q = 0; q < 6; ++q) { double av = fabs(dT[q]); if (av > mx) mx = av; }
      if (mx <= 1e-13) break;                                           // :786
    }
  }

  double* po = a.pose_io + (size_t)f * 16;
  if (ok) {
    for (int e = 0; e < 12; ++e) po[e] = T[e];
    po[12] = 0; po[13] = 0; po[14] = 0; po[15] = 1;
  }
  if (ran_gn && a.cov) {
    double cov[36];
    inverse6(A, cov);                                                   // :790
    for (int e = 0; e < 36; ++e) a.cov[(size_t)f * 36 + e] = cov[e];
  }
  if (a.ok) a.ok[f] = ok;
  if (a.iters) a.iters[f] = iters;
  if (a.updated) a.updated[f] = (ok && ran_gn) ? 1 : 0;
  if (a.mode == 1) {                       // tracking loop: route the stream to optimisePose or to initialise()
    if (ok) { if (a.set_gn_if_ok) a.set_gn_if_ok[f] = 1; }
    else if (a.set_init_if_fail) a.set_init_if_fail[f] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// K3c — gauss_newton_kernel: PoseEstimator::optimisePose (pose_estimator.cpp:733-792) with a GROUP of G lanes per frame.
// The thread-per-frame version above is a ~100 us serial chain per frame (5 iterations x [5 Jacobians with 10 divisions each,
// a pivoted 6x6 LDL^T with 56 divisions, an exponential map]) and keeps fewer than two warps per SM busy at 8192 frames.
// Here, per iteration:
//   1. lane j computes project2d, the residual and the two Jacobian rows of correspondence j       (:759-772, :932-960)
//   2. lane e owns entry e of (A, b) — 21 unique entries of the symmetric A, 6 of b — and sums the contributions of the
//      correspondences IN CORRESPONDENCE ORDER, i.e. exactly the reference's  A += J^T J,  b += J^T e  (:774-775)
//   3. the pivoted LDL^T (:778) runs with the Schur updates of one elimination step spread over the lanes; row/column
//      exchanges are a permutation held in a register, so no data moves.  Every entry sees the same operations in the same
//      order as oracle ldlt_solve6 / the serial ldlt_solve6 above, hence identical bits and identical iteration counts
//   4. exponential map and pose update (:781): the 9 + 9 entries of R and V over the lanes
// and at the end the covariance A^-1 (:790) by the same Gauss-Jordan as inverse6 with one tableau column per lane.
// G = 8 (four frames per warp) for large batches, G = 32 when few frames are in flight (single-camera latency).
// ------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 64;
constexpr int kGnCooperativeMaxFrames = 1024;

struct GnScratch {
  double J[MPE_MAX_LEDS][14];   // per correspondence: J0[6], J1[6], e0, e1
  double A0[36];                // normal matrix of the current iteration (kept: the covariance is its inverse)
  double Aw[36];                // working copy, factorised in place
  double b[6], dT[6];
  double T[12];                 // pose, top three rows, row-major
  double rv[18];                // R[9], V[9] of the exponential map
  double tab[72];               // 6 x 12 Gauss-Jordan tableau
  int used[MPE_MAX_LEDS];
};

__device__ __forceinline__ int nib(uint32_t pk, int i) { return (int)((pk >> (4 * i)) & 15u); }
__device__ __forceinline__ uint32_t nib_swap(uint32_t pk, int i, int j) {
  const uint32_t a = (pk >> (4 * i)) & 15u, b = (pk >> (4 * j)) & 15u;
  pk &= ~((15u << (4 * i)) | (15u << (4 * j)));
  return pk | (b << (4 * i)) | (a << (4 * j));
}

template <int G>
__global__ void __launch_bounds__(kGnThreads) gauss_newton_kernel(const K3Args a, int gate_on_ok) {
  pdl_enter();
  constexpr int kGroups = kGnThreads / G;
  __shared__ GnScratch scratch[kGroups];
  const int gid = threadIdx.x / G, gl = threadIdx.x % G;
  const int f = blockIdx.x * kGroups + gid;
  // Control flow is WARP-uniform: the groups of a warp run the stages in lock-step and meet at full-mask __syncwarp()s (a
  // group that only synchronised with its own lanes would never reconverge with its neighbours after the first data-dependent
  // branch, and the warp would execute its groups one after the other — measured: 4x slower with G = 8).  A group without
  // work (no frame, inactive, check failed) or one that has converged keeps walking through the stages with `run` false.
  bool alive = f < a.n_frames;
  if (alive && a.active && !a.active[f]) alive = false;
  if (alive && gate_on_ok && !a.ok[f]) alive = false;              // checkCorrespondences failed: no optimisePose
  GnScratch& S = scratch[gid];
  const double* K = a.cam.K;
  const double* mk = a.pp.markers;
  int k = 0;
  const double* det = a.det;
  const uint32_t* corr = a.corr;
  if (alive) {
    const int n_det = a.n_det[f];
    k = a.n_corr[f];
    if (n_det < 0 || n_det > MPE_MAX_DET) k = 0;
    if (k > MPE_MAX_LEDS) k = MPE_MAX_LEDS;
    det = a.det + (size_t)f * a.det_stride * 2;
    corr = a.corr + (size_t)f * 2 * MPE_MAX_LEDS;
    for (int e = gl; e < 12; e += G) S.T[e] = a.pose_io[(size_t)f * 16 + e];
    for (int e = gl; e < 36; e += G) S.A0[e] = 0;
  }
  const double fx = K[0], fy = K[4];
  __syncwarp();

  int iters = 0;
  bool done = false;
  for (int it = 0; it < 500; ++it) {                               // max_itr :738
    const bool run = alive && !done;
    if (!__any_sync(0xffffffffu, run)) break;
    double T[12], KT[12];
    if (run) {
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = S.T[e];
      kt_product(K, T, KT);                                        // project2d :251-268 (same KT for every point)
      // ---- 1. one correspondence per lane
      for (int j = gl; j < k; j += G) {
        const uint32_t led1 = corr[2 * j], det1 = corr[2 * j + 1];
        S.used[j] = (det1 != 0);                                   // :761
        if (det1 == 0) continue;
        const int led = (int)led1 - 1, di = (int)det1 - 1;
        const double ox = mk[3 * led], oy = mk[3 * led + 1], oz = mk[3 * led + 2];
        double pu, pv;
        kt_project(KT, ox, oy, oz, pu, pv);
        const double e0 = det[2 * di] - pu, e1 = det[2 * di + 1] - pv;   // :769
        // computeJacobian :932-960
        const double x = T[0] * ox + T[1] * oy + T[2] * oz + T[3] * 1.0;
        const double y = T[4] * ox + T[5] * oy + T[6] * oz + T[7] * 1.0;
        const double z = T[8] * ox + T[9] * oy + T[10] * oz + T[11] * 1.0;
        const double z_2 = z * z;
        double* Jj = S.J[j];
        Jj[0] = 1 / z * fx; Jj[1] = 0; Jj[2] = -x / z_2 * fx; Jj[3] = -x * y / z_2 * fx; Jj[4] = (1 + (x * x / z_2)) * fx; Jj[5] = -y / z * fx;
        Jj[6] = 0; Jj[7] = 1 / z * fy; Jj[8] = -y / z_2 * fy; Jj[9] = -(1 + y * y / z_2) * fy; Jj[10] = x * y / z_2 * fy; Jj[11] = x / z * fy;
        Jj[12] = e0; Jj[13] = e1;
      }
    }
    __syncwarp();
    // ---- 2. one entry of (A, b) per lane, contributions added in correspondence order
    if (run) {
      for (int e = gl; e < 27; e += G) {
        double acc = 0;
        if (e < 21) {
          int r = 0, c = e;                                        // e -> (r, c), c <= r, rows of the lower triangle one after another
          while (c > r) { c -= r + 1; ++r; }
          for (int j = 0; j < k; ++j) {
            if (!S.used[j]) continue;
            const double* Jj = S.J[j];
            acc += Jj[r] * Jj[c] + Jj[6 + r] * Jj[6 + c];
          }
          S.A0[r * 6 + c] = acc; S.A0[c * 6 + r] = acc;            // J0[r]*J0[c] == J0[c]*J0[r]: the reference's two entries are equal
        } else {
          const int r = e - 21;
          for (int j = 0; j < k; ++j) {
            if (!S.used[j]) continue;
            const double* Jj = S.J[j];
            acc += Jj[r] * Jj[12] + Jj[6 + r] * Jj[13];
          }
          S.b[r] = acc;
        }
      }
    }
    __syncwarp();
    // ---- 3. dT = A.ldlt().solve(b)  (:778), arithmetic of ldlt_solve6: left-looking LDL^T of P A P^T.  Entries that have not
    //         been reached yet are the original ones, so the exchanges are a permutation of labels (nibbles of pk) and no
    //         data moves: L(i,j) = Aw[perm[i]][j], untouched entries = A0[perm[i]][perm[j]].
    uint32_t pk = 0x543210u;                                       // perm[i] = nibble i
    double Dg[6] = {1, 1, 1, 1, 1, 1};
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) {
      if (run) {
        int p = kk;
        double best = fabs(S.A0[nib(pk, kk) * 7]);
#pragma unroll
        for (int i = kk + 1; i < 6; ++i) {
          const double v = fabs(S.A0[nib(pk, i) * 7]);
          if (v > best) { best = v; p = i; }
        }
        pk = nib_swap(pk, kk, p);
        const int pkk = nib(pk, kk);
        double temp[6] = {0, 0, 0, 0, 0, 0};
        double dk = S.A0[pkk * 7];
        if (kk > 0) {
#pragma unroll
          for (int j = 0; j < kk; ++j) temp[j] = Dg[j] * S.Aw[pkk * 6 + j];
          double sacc = S.Aw[pkk * 6] * temp[0];
#pragma unroll
          for (int j = 1; j < kk; ++j) sacc += S.Aw[pkk * 6 + j] * temp[j];
          dk -= sacc;
        }
        Dg[kk] = dk;
        for (int i = kk + 1 + gl; i < 6; i += G) {                  // column kk of L, one row per lane
          const int pi = nib(pk, i);
          double v = S.A0[pi * 6 + pkk];
          if (kk > 0) {
            double t = S.Aw[pi * 6] * temp[0];
#pragma unroll
            for (int j = 1; j < kk; ++j) t += S.Aw[pi * 6 + j] * temp[j];
            v -= t;
          }
          if (fabs(dk) > 0.0) v /= dk;
          S.Aw[pi * 6 + kk] = v;
        }
      }
      __syncwarp();
    }
    double dT[6] = {0, 0, 0, 0, 0, 0};
    if (run) {
      double y[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) y[i] = S.b[nib(pk, i)];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] -= S.Aw[nib(pk, i) * 6 + j] * y[j];
#pragma unroll
      for (int i = 0; i < 6; ++i) { if (fabs(Dg[i]) > 2.2250738585072014e-308) y[i] /= Dg[i]; else y[i] = 0.0; }
#pragma unroll
      for (int i = 5; i >= 0; --i)
#pragma unroll
        for (int j = 5; j > i; --j) y[i] -= S.Aw[nib(pk, j) * 6 + i] * y[j];
      if (gl == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) S.dT[nib(pk, i)] = y[i];
      }
    }
    __syncwarp();
    // ---- 4. T <- exp(dT) * T  (:781, :962-994), arithmetic of exp_map_left_multiply
    double ux = 0, uy = 0, uz = 0;
    if (run) {
#pragma unroll
      for (int i = 0; i < 6; ++i) dT[i] = S.dT[i];
      ux = dT[0]; uy = dT[1]; uz = dT[2];
      const double wx = dT[3], wy = dT[4], wz = dT[5];
      const double theta = sqrt(wx * wx + wy * wy + wz * wz);
      const double theta_squared = theta * theta;
      double sn = 0, cs = 0, kv1 = 0, kv2 = 0;
      if (theta != 0) {
        sn = sin(theta); cs = cos(theta);
        kv1 = (1 - cs) / (theta_squared); kv2 = (theta - sn) / (theta_squared * theta);
      }
      for (int e = gl; e < 9; e += G) {
        const int i = e / 3, j = e - 3 * i;
        // row i and column j of O = [w]x (selects, no run-time indexed local arrays)
        const double Oi0 = i == 0 ? 0.0 : (i == 1 ? wz : -wy), Oi1 = i == 0 ? -wz : (i == 1 ? 0.0 : wx), Oi2 = i == 0 ? wy : (i == 1 ? -wx : 0.0);
        const double Oj0 = j == 0 ? 0.0 : (j == 1 ? -wz : wy), Oj1 = j == 0 ? wz : (j == 1 ? 0.0 : -wx), Oj2 = j == 0 ? -wy : (j == 1 ? wx : 0.0);
        const double Oij = j == 0 ? Oi0 : (j == 1 ? Oi1 : Oi2);
        const double O2ij = Oi0 * Oj0 + Oi1 * Oj1 + Oi2 * Oj2;
        const double I = (i == j) ? 1.0 : 0.0;
        double r_ij = I, v_ij = I;
        if (theta != 0) {
          r_ij = I + Oij / theta * sn + O2ij / theta_squared * (1 - cs);
          v_ij = I + kv1 * Oij + kv2 * O2ij;
        }
        S.rv[e] = r_ij; S.rv[9 + e] = v_ij;
      }
    }
    __syncwarp();
    if (run) {
      for (int e = gl; e < 12; e += G) {
        const int i = e >> 2, j = e & 3;
        const double ti = S.rv[9 + 3 * i] * ux + S.rv[9 + 3 * i + 1] * uy + S.rv[9 + 3 * i + 2] * uz;
        const double bottom = (j == 3) ? 1.0 : 0.0;
        const double T0j = j == 0 ? T[0] : (j == 1 ? T[1] : (j == 2 ? T[2] : T[3]));
        const double T1j = j == 0 ? T[4] : (j == 1 ? T[5] : (j == 2 ? T[6] : T[7]));
        const double T2j = j == 0 ? T[8] : (j == 1 ? T[9] : (j == 2 ? T[10] : T[11]));
        double sacc = S.rv[3 * i] * T0j;
        sacc += S.rv[3 * i + 1] * T1j;
        sacc += S.rv[3 * i + 2] * T2j;
        sacc += ti * bottom;
        S.T[e] = sacc;
      }
    }
    __syncwarp();
    if (run) {
      ++iters;
      double mx = -1;                                               // norm_max :1073-1085
#pragma unroll
      for (int q = 0; q < 6; ++q) { const double av = fabs(dT[q]); if (av > mx) mx = av; }
      if (mx <= 1e-13) done = true;                                 // :786
    }
  }

  // ---- covariance = A^-1 of the last iteration (:790), arithmetic of inverse6; the row exchanges are a permutation
  if (alive)
    for (int e = gl; e < 72; e += G) { const int i = e / 12, j = e - 12 * i; S.tab[e] = (j < 6) ? S.A0[i * 6 + j] : ((j - 6 == i) ? 1.0 : 0.0); }
  __syncwarp();
  uint32_t rp = 0x543210u;
#pragma unroll
  for (int kk = 0; kk < 6; ++kk) {
    int rk = 0;
    double pv = 1.0;
    if (alive) {
      int p = kk;
      double best = fabs(S.tab[nib(rp, kk) * 12 + kk]);
#pragma unroll
      for (int i = kk + 1; i < 6; ++i) {
        const double v = fabs(S.tab[nib(rp, i) * 12 + kk]);
        if (v > best) { best = v; p = i; }
      }
      rp = nib_swap(rp, kk, p);
      rk = nib(rp, kk);
      pv = S.tab[rk * 12 + kk];
    }
    __syncwarp();                                                   // everybody has read the pivot before the row is scaled
    if (alive)
      for (int j = gl; j < 12; j += G) S.tab[rk * 12 + j] = S.tab[rk * 12 + j] / pv;
    __syncwarp();
    double fct[6] = {0, 0, 0, 0, 0, 0};
    if (alive) {
#pragma unroll
      for (int i = 0; i < 6; ++i) fct[i] = S.tab[nib(rp, i) * 12 + kk];
    }
    __syncwarp();                                                   // factors read before column kk is overwritten
    if (alive) {
      for (int e = gl; e < 72; e += G) {
        const int i = e / 12, j = e - 12 * i;
        if (i == kk) continue;
        const double fi = (i == 0) ? fct[0] : (i == 1) ? fct[1] : (i == 2) ? fct[2] : (i == 3) ? fct[3] : (i == 4) ? fct[4] : fct[5];
        if (fi != 0) { const int ri = nib(rp, i); S.tab[ri * 12 + j] = S.tab[ri * 12 + j] - fi * S.tab[rk * 12 + j]; }
      }
    }
    __syncwarp();
  }
  if (!alive) return;
  if (a.cov)
    for (int e = gl; e < 36; e += G) { const int i = e / 6, j = e - 6 * i; a.cov[(size_t)f * 36 + e] = S.tab[nib(rp, i) * 12 + 6 + j]; }
  double* po = a.pose_io + (size_t)f * 16;
  for (int e = gl; e < 12; e += G) po[e] = S.T[e];
  if (gl == 0) {
    po[12] = 0; po[13] = 0; po[14] = 0; po[15] = 1;
    if (a.ok) a.ok[f] = 1;
    if (a.iters) a.iters[f] = iters;
    if (a.updated) a.updated[f] = 1;
  }
}

static bool wide_ok() {      // MPE_K3_WIDE=0 keeps the thread-per-subset check kernel for every batch size
  static int v = -1;
  if (v < 0) { const char* e = getenv("MPE_K3_WIDE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

cudaError_t launch_validate_refine(const K3Args& a, cudaStream_t st) {
  if ((a.mode == 0 || a.mode == 1) && a.n_frames <= kGnCooperativeMaxFrames && wide_ok()) {
    const int n_obj = a.pp.n_obj;
    size_t smem = sizeof(CheckFrame) + 4 * 32 * (sizeof(double) + sizeof(int)) + (size_t)32 * n_obj * 3 * sizeof(double) + 32 * sizeof(int);
    cudaError_t e = launch_k(check_wide_kernel, a.n_frames, kWideThreads, smem, st, a);
    if (e != cudaSuccess) return e;
  } else if (a.mode == 0 || a.mode == 1) {
    const int n_obj = a.pp.n_obj;
    int Nmax = (n_obj >= 4) ? n_obj * (n_obj - 1) * (n_obj - 2) / 6 : 1;
    int G = (Nmax <= kK3aThreads) ? kK3aThreads / Nmax : 1;
    if (G > 32) G = 32;
    size_t smem = (size_t)G * sizeof(CheckFrame) + (size_t)kK3aThreads * n_obj * 3 * sizeof(double) + kK3aThreads * sizeof(int);
    static SmemAttrCache configured;
    {
      cudaError_t e = ensure_dynamic_smem(check_kernel, smem, configured);
      if (e != cudaSuccess) return e;
    }
    int grid = (a.n_frames + G - 1) / G;
    cudaError_t e = launch_k(check_kernel, grid, kK3aThreads, smem, st, a, G, Nmax);
    if (e != cudaSuccess) return e;
  }
  // Gauss-Newton layout.  Measured on B200: one frame takes 105 us with a thread per frame and 36.5 us with 32 lanes per frame,
  // but at 8192 frames the thread-per-frame kernel (every lane busy, 0.25 ms) beats 8 or 32 lanes per frame (0.22 / 0.27 ms plus
  // the separate Kabsch launch) — so the lane-cooperative kernel serves small batches (single cameras), the serial one large ones.
  // MPE_K3_GN overrides: 0 = always thread per frame, 8 / 32 = always that many lanes per frame.
  static int gn_mode = -1;
  if (gn_mode < 0) { const char* e = getenv("MPE_K3_GN"); gn_mode = e ? atoi(e) : -2; }
  const bool cooperative = (gn_mode == 8 || gn_mode == 32) || (gn_mode != 0 && a.n_frames <= kGnCooperativeMaxFrames);
  if (!cooperative || a.mode == 1) {
    return launch_k(refine_kernel, (a.n_frames + kK3bThreads - 1) / kK3bThreads, kK3bThreads, 0, st, a);
  }
  // mode 0: acceptance test + Kabsch with a thread per frame (refine_kernel in check-only mode), then the lane-cooperative
  // Gauss-Newton for the frames it accepted; mode 2: Gauss-Newton only
  if (a.mode == 0) {
    K3Args chk = a;
    chk.mode = 1;
    cudaError_t e = launch_k(refine_kernel, (a.n_frames + kK3bThreads - 1) / kK3bThreads, kK3bThreads, 0, st, chk);
    if (e != cudaSuccess) return e;
  }
  const int gate = (a.mode == 0) ? 1 : 0;
  const int lanes = (gn_mode == 8) ? 8 : 32;
  if (lanes == 32) return launch_k(gauss_newton_kernel<32>, (a.n_frames + 1) / 2, kGnThreads, 0, st, a, gate);
  return launch_k(gauss_newton_kernel<8>, (a.n_frames + 7) / 8, kGnThreads, 0, st, a, gate);
}

}  // namespace mpe
