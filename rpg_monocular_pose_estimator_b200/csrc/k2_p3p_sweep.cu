// K2 — PoseEstimator::initialise (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:544-721):
// brute-force correspondence search.  For every 3-subset of the detections (lexicographic,
// Combinations::combinationsNoReplacement) and every ordered 3-tuple of LEDs
// (Combinations::permutationsNoReplacement) one Kneip P3P problem is solved, each of its four solutions is
// scored by back-projecting the unused LEDs and voting into the n_det x n_obj integer histogram.  The
// histogram is then decoded by PoseEstimator::correspondencesFromHistogram (:344-370).
//
// Mapping: one THREAD per P3P problem (the whole solve + the four scorings are straight-line scalar FP64
// code, so a warp per problem would idle 31 lanes; with thousands of frames per launch there is no shortage
// of parallelism).  A CTA accumulates votes in a shared-memory histogram; `split` CTAs per frame merge into
// the global histogram with integer atomics (order-independent, hence exact) and the last CTA to finish a
// frame decodes the correspondences (block-wide argmax = a 256-entry scan by one warp).
//
// FP64 / latency bound, not HBM bound: it reads < 1 KB per frame.  Compiled with -fmad=false.
#include "mpe_internal.cuh"
#include "p3p_device.cuh"

namespace mpe {

#ifndef MPE_K2_THREADS
#define MPE_K2_THREADS 256
#endif
#ifndef MPE_K2_MINBLOCKS
#define MPE_K2_MINBLOCKS 3
#endif
constexpr int kK2Threads = MPE_K2_THREADS;

// lexicographic unranking of a 3-combination of {0..n-1}
__device__ __forceinline__ void unrank_comb3(int n, int idx, int& a, int& b, int& c) {
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

// row j of Combinations::permutationsNoReplacement(n,3) (combinations.cpp:127-244), 0-based:
// block j/6 = lexicographic combination (a<b<c); rows [c b a],[c a b],[b c a],[b a c],[a b c],[a c b]
__device__ __forceinline__ void unrank_perm3(int n, int j, int& p0, int& p1, int& p2) {
  int a, b, c;
  unrank_comb3(n, j / 6, a, b, c);
  switch (j % 6) {
    case 0: p0 = c; p1 = b; p2 = a; break;
    case 1: p0 = c; p1 = a; p2 = b; break;
    case 2: p0 = b; p1 = c; p2 = a; break;
    case 3: p0 = b; p1 = a; p2 = c; break;
    case 4: p0 = a; p1 = b; p2 = c; break;
    default: p0 = a; p1 = c; p2 = b; break;
  }
}

struct K2Shared {
  double det[MPE_MAX_DET][2];
  double bearing[MPE_MAX_DET][3];
  double marker[MPE_MAX_LEDS][3];
  uint32_t hist[MPE_MAX_DET * MPE_MAX_LEDS];
  int is_last;
};

// PoseEstimator::calculateImageVectors (pose_estimator.cpp:288-301)
__device__ __forceinline__ void bearing_vector(const DevCamera& cam, double u, double v, double out[3]) {
  double x = (u - cam.K[2]) / cam.K[0];
  double y = (v - cam.K[5]) / cam.K[4];
  double z = 1;
  double n = sqrt(x * x + y * y + z * z);
  out[0] = x / n; out[1] = y / n; out[2] = z / n;
}

// PoseEstimator::correspondencesFromHistogram (pose_estimator.cpp:344-370) on hist (row = detection,
// col = LED, row-major n_det x n_obj).  Eigen's maxCoeff(&r,&c) visits column-major and keeps the first
// strict maximum; only the chosen column is cleared.  Executed by one thread (<= 256 entries, <= 16 rounds).
__device__ int decode_histogram(uint32_t* hist, int n_det, int n_obj, uint32_t threshold, uint32_t* corr) {
  int n = 0;
  for (int j = 0; j < n_obj; ++j) {
    uint32_t mv = hist[0];
    int ri = 0, ci = 0;
    for (int c = 0; c < n_obj; ++c)
      for (int r = 0; r < n_det; ++r) {
        uint32_t v = hist[r * n_obj + c];
        if (v > mv) { mv = v; ri = r; ci = c; }
      }
    if (mv < threshold) break;
    corr[2 * n] = (uint32_t)ci + 1;
    corr[2 * n + 1] = (uint32_t)ri + 1;
    ++n;
    for (int r = 0; r < n_det; ++r) hist[r * n_obj + ci] = 0;
  }
  return n;
}

__global__ void __launch_bounds__(kK2Threads, MPE_K2_MINBLOCKS) p3p_sweep_kernel(const K2Args a) {
  __shared__ K2Shared sh;
  const int f = blockIdx.x / a.split;
  const int part = blockIdx.x - f * a.split;
  const int tid = threadIdx.x;
  if (a.active && !a.active[f]) return;
  const int n_det = a.n_det[f];
  const int n_obj = a.pp.n_obj;
  if (n_det < 4 || n_det > MPE_MAX_DET) {      // pose_estimator.cpp:80 (min_num_leds_detected_ = 4)
    if (part == 0 && tid == 0) {
      a.n_corr[f] = 0;
      if (n_det > MPE_MAX_DET) a.frame_flags[f] |= MPE_F_TOO_MANY_DET;
    }
    return;
  }
  const double* det = a.det + (size_t)f * a.det_stride * 2;
  if (tid < n_det) {
    double u = det[2 * tid], v = det[2 * tid + 1];
    sh.det[tid][0] = u; sh.det[tid][1] = v;
    bearing_vector(a.cam, u, v, sh.bearing[tid]);
  }
  if (tid < n_obj) {
    sh.marker[tid][0] = a.pp.markers[3 * tid];
    sh.marker[tid][1] = a.pp.markers[3 * tid + 1];
    sh.marker[tid][2] = a.pp.markers[3 * tid + 2];
  }
  for (int i = tid; i < MPE_MAX_DET * MPE_MAX_LEDS; i += kK2Threads) sh.hist[i] = 0;
  __syncthreads();

  const int n_comb = n_det * (n_det - 1) * (n_det - 2) / 6;
  const int n_perm = n_obj * (n_obj - 1) * (n_obj - 2);
  const int total = n_comb * n_perm;
  const double tol_sq_max = a.pp.back_proj_sq_max;
  const int n_unused_obj = n_obj - 3;

  for (int t = part * kK2Threads + tid; t < total; t += a.split * kK2Threads) {
    const int ci = t / n_perm, pj = t - ci * n_perm;
    int d0, d1, d2, o0, o1, o2;
    unrank_comb3(n_det, ci, d0, d1, d2);
    unrank_perm3(n_obj, pj, o0, o1, o2);

    P3PSetup S;
    int rc = p3p_setup(v_make(sh.bearing[d0][0], sh.bearing[d0][1], sh.bearing[d0][2]),
                       v_make(sh.bearing[d1][0], sh.bearing[d1][1], sh.bearing[d1][2]),
                       v_make(sh.bearing[d2][0], sh.bearing[d2][1], sh.bearing[d2][2]),
                       v_make(sh.marker[o0][0], sh.marker[o0][1], sh.marker[o0][2]),
                       v_make(sh.marker[o1][0], sh.marker[o1][1], sh.marker[o1][2]),
                       v_make(sh.marker[o2][0], sh.marker[o2][1], sh.marker[o2][2]), S);
    if (rc != 0) continue;

    for (int k = 0; k < 4; ++k) {
      double H[12];
      if (!p3p_solution(S, k, H)) continue;
      if (!h_is_finite(H)) continue;                       // pose_estimator.cpp:653
      double Hi[12], KT[12];
      h_inverse(H, Hi);                                    // :660
      kt_product(a.cam.K, Hi, KT);
      // back-project the unused LEDs (:658-661)
      double bu[MPE_MAX_LEDS - 3], bv[MPE_MAX_LEDS - 3];
      int m = 0;
      for (int ll = 0; ll < n_obj; ++ll) {
        if (ll == o0 || ll == o1 || ll == o2) continue;
        kt_project(KT, sh.marker[ll][0], sh.marker[ll][1], sh.marker[ll][2], bu[m], bv[m]);
        ++m;
      }
      // nearest back-projection for every unused detection (:664, calculateMinDistancesAndPairs :862-906)
      uint32_t within = 0;        // bit i: unused detection i is within tolerance
      unsigned long long pairs = 0;   // 4 bits per unused detection: index of the nearest unused LED
      int ui = 0;
      for (int kk = 0; kk < n_det; ++kk) {
        if (kk == d0 || kk == d1 || kk == d2) continue;
        double best = HUGE_VAL;
        int bj = 0;
        for (int j = 0; j < n_unused_obj; ++j) {
          double dx = sh.det[kk][0] - bu[j], dy = sh.det[kk][1] - bv[j];
          double d2v = dx * dx + dy * dy;
          if (d2v < best) { best = d2v; bj = j; }
        }
        if (best <= tol_sq_max) within |= 1u << ui;        // :671  sqrt(best) < tol, see DevPoseParams::back_proj_sq_max
        pairs |= (unsigned long long)bj << (4 * ui);
        ++ui;
      }
      if (within) {                                         // :676
        atomicAdd(&sh.hist[d0 * n_obj + o0], 1u);           // :680-685
        atomicAdd(&sh.hist[d1 * n_obj + o1], 1u);
        atomicAdd(&sh.hist[d2 * n_obj + o2], 1u);
        ui = 0;
        for (int kk = 0; kk < n_det; ++kk) {                // :687-695
          if (kk == d0 || kk == d1 || kk == d2) continue;
          if (within & (1u << ui)) {
            int bj = (int)((pairs >> (4 * ui)) & 0xf);
            // bj-th unused LED -> LED index
            int obj = -1, cnt = 0;
            for (int ll = 0; ll < n_obj; ++ll) {
              if (ll == o0 || ll == o1 || ll == o2) continue;
              if (cnt == bj) { obj = ll; break; }
              ++cnt;
            }
            atomicAdd(&sh.hist[kk * n_obj + obj], 1u);
          }
          ++ui;
        }
      }
    }
  }
  __syncthreads();

  uint32_t* ghist = a.hist + (size_t)f * MPE_MAX_DET * MPE_MAX_LEDS;
  if (a.split > 1) {
    for (int i = tid; i < n_det * n_obj; i += kK2Threads)
      if (sh.hist[i]) atomicAdd(&ghist[i], sh.hist[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      uint32_t prev = atomicAdd(&a.done_counter[f], 1u);
      sh.is_last = (prev == (uint32_t)a.split - 1);
    }
    __syncthreads();
    if (!sh.is_last) return;
    __threadfence();
    for (int i = tid; i < n_det * n_obj; i += kK2Threads) sh.hist[i] = __ldcg(&ghist[i]);
    __syncthreads();
  } else {
    for (int i = tid; i < n_det * n_obj; i += kK2Threads) ghist[i] = sh.hist[i];
  }
  if (tid == 0) {
    // pose_estimator.cpp:704: decode only if the histogram is not all zero
    bool all_zero = true;
    for (int i = 0; i < n_det * n_obj; ++i) all_zero = all_zero && (sh.hist[i] == 0);
    int n = 0;
    if (!all_zero) n = decode_histogram(sh.hist, n_det, n_obj, a.pp.histogram_threshold, a.corr + (size_t)f * 2 * MPE_MAX_LEDS);
    a.n_corr[f] = n;
    a.frame_flags[f] |= MPE_F_INITIALISED;
  }
}

cudaError_t launch_p3p_sweep(const K2Args& a, cudaStream_t st) {
  p3p_sweep_kernel<<<a.n_frames * a.split, kK2Threads, 0, st>>>(a);
  return cudaGetLastError();
}

// ---- stand-alone P3P (mpe_p3p_compute_poses): one thread per problem ----
__global__ void p3p_batch_kernel(const double* __restrict__ f, const double* __restrict__ P, int n, double* __restrict__ sol, int* __restrict__ status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* fi = f + 9 * (size_t)i;
  const double* Pi = P + 9 * (size_t)i;
  P3PSetup S;
  int rc = p3p_setup(v_make(fi[0], fi[1], fi[2]), v_make(fi[3], fi[4], fi[5]), v_make(fi[6], fi[7], fi[8]),
                     v_make(Pi[0], Pi[1], Pi[2]), v_make(Pi[3], Pi[4], Pi[5]), v_make(Pi[6], Pi[7], Pi[8]), S);
  status[i] = rc;
  double* out = sol + 48 * (size_t)i;
  if (rc != 0) {
    for (int k = 0; k < 48; ++k) out[k] = 0.0;
    return;
  }
  for (int k = 0; k < 4; ++k) {
    double H[12];
    if (!p3p_solution(S, k, H)) {
      for (int e = 0; e < 12; ++e) H[e] = nan("");          // what the full evaluation yields: NaN in every entry of R (C may differ)
    }
    for (int e = 0; e < 12; ++e) out[12 * k + e] = H[e];
  }
}

cudaError_t launch_p3p_batch(const double* f, const double* P, int n, double* sol, int* status, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  p3p_batch_kernel<<<(n + 127) / 128, 128, 0, st>>>(f, P, n, sol, status);
  return cudaGetLastError();
}

}  // namespace mpe
