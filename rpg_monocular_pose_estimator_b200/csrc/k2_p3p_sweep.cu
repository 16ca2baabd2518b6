// K2 — PoseEstimator::initialise (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:544-721):
// brute-force correspondence search.  For every 3-subset of the detections (lexicographic,
// Combinations::combinationsNoReplacement) and every ordered 3-tuple of LEDs
// (Combinations::permutationsNoReplacement) one Kneip P3P problem is solved, each of its four solutions is
// scored by back-projecting the unused LEDs and voting into the n_det x n_obj integer histogram.  The
// histogram is then decoded by PoseEstimator::correspondencesFromHistogram (:344-370).
//
// Mapping: one THREAD per P3P problem (the whole solve + the four scorings are straight-line scalar FP64
// code, so a warp per problem would idle 31 lanes; with thousands of frames per launch there is no shortage
// of parallelism).  Votes are integer atomics on the frame's global histogram (order-independent, hence exact);
// bearings come from a tiny prologue kernel, the histogram is decoded by a one-thread-per-frame epilogue kernel.
//
// FP64 / latency bound, not HBM bound: it reads < 1 KB per frame.  Compiled with -fmad=false.
#include "mpe_internal.cuh"
#include "p3p_device.cuh"
#include <cstdlib>

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

// PoseEstimator::calculateImageVectors (pose_estimator.cpp:288-301)
__device__ __forceinline__ void bearing_vector(const DevCamera& cam, double u, double v, double out[3]) {
  double x = (u - cam.K[2]) / cam.K[0];
  double y = (v - cam.K[5]) / cam.K[4];
  double z = 1;
  double n = sqrt(x * x + y * y + z * z);
  out[0] = x / n; out[1] = y / n; out[2] = z / n;
}

// setImagePoints -> image_vectors_: one thread per (frame, detection)
__global__ void bearings_kernel(const K2Args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = i / MPE_MAX_DET, d = i - f * MPE_MAX_DET;
  if (f >= a.n_frames) return;
  if (a.active && !a.active[f]) return;
  const int n_det = a.n_det[f];
  if (d >= n_det || n_det > MPE_MAX_DET) return;
  const double* det = a.det + (size_t)f * a.det_stride * 2;
  bearing_vector(a.cam, det[2 * d], det[2 * d + 1], a.bearings + ((size_t)f * MPE_MAX_DET + d) * 3);
}

// PoseEstimator::correspondencesFromHistogram (pose_estimator.cpp:344-370) on hist (row = detection,
// col = LED, row-major n_det x n_obj).  Eigen's maxCoeff(&r,&c) visits column-major and keeps the first
// strict maximum; only the chosen column is cleared.
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

// The sweep.  A CTA walks over the frames blockIdx.x, blockIdx.x + gridDim.x, ... and treats their P3P problems as ONE flat
// sequence dealt round-robin to its threads (`rot` carries the position over from frame to frame), so no thread waits for a
// partially filled last iteration and there is no block barrier at all: votes go straight to the global histogram with
// integer atomics (about a hundred per frame).  (The first version kept a shared-memory histogram per CTA and synchronised
// twice per frame; 29 % of its warp stalls were barrier waits, 600 problems over 256 threads leave a 35 %-filled third pass.)
__global__ void __launch_bounds__(kK2Threads, MPE_K2_MINBLOCKS) p3p_sweep_kernel(const K2Args a) {
  const int tid = threadIdx.x;
  const int n_obj = a.pp.n_obj;
  const int n_perm = n_obj * (n_obj - 1) * (n_obj - 2);
  const int n_unused_obj = n_obj - 3;
  const double tol_sq_max = a.pp.back_proj_sq_max;
  // marker coordinates in shared memory: lanes of a warp index them with different permutations, which the constant
  // bank (kernel parameters) would serialise
  __shared__ double mk[3 * MPE_MAX_LEDS];
  if (tid < 3 * MPE_MAX_LEDS) mk[tid] = a.pp.markers[tid];
  __syncthreads();                                       // the only barrier of the kernel
  const int n_units = a.n_frames * a.split;
  int rot = 0;                                          // flat position (mod blockDim) where this unit's first problem falls
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int f = unit / a.split, part = unit - f * a.split;
    if (a.active && !a.active[f]) continue;
    const int n_det = a.n_det[f];
    if (n_det < 4 || n_det > MPE_MAX_DET) continue;     // pose_estimator.cpp:80 (min_num_leds_detected_ = 4); flagged by decode_kernel
    const int n_comb = n_det * (n_det - 1) * (n_det - 2) / 6;
    const int total = n_comb * n_perm;
    const int chunk = (total + a.split - 1) / a.split;
    const int t_begin = part * chunk, t_end = min(total, t_begin + chunk);
    const int count = max(t_end - t_begin, 0);
    const double* det = a.det + (size_t)f * a.det_stride * 2;
    const double* bear = a.bearings + (size_t)f * MPE_MAX_DET * 3;
    uint32_t* ghist = a.hist + (size_t)f * MPE_MAX_DET * MPE_MAX_LEDS;

    int first = tid - rot;
    if (first < 0) first += kK2Threads;
    for (int t = t_begin + first; t < t_end; t += kK2Threads) {
      const int ci = t / n_perm, pj = t - ci * n_perm;
      int d0, d1, d2, o0, o1, o2;
      unrank_comb3(n_det, ci, d0, d1, d2);
      unrank_perm3(n_obj, pj, o0, o1, o2);

      P3PSetup S;
      int rc = p3p_setup(v_make(bear[3 * d0], bear[3 * d0 + 1], bear[3 * d0 + 2]), v_make(bear[3 * d1], bear[3 * d1 + 1], bear[3 * d1 + 2]),
                         v_make(bear[3 * d2], bear[3 * d2 + 1], bear[3 * d2 + 2]), v_make(mk[3 * o0], mk[3 * o0 + 1], mk[3 * o0 + 2]),
                         v_make(mk[3 * o1], mk[3 * o1 + 1], mk[3 * o1 + 2]), v_make(mk[3 * o2], mk[3 * o2 + 1], mk[3 * o2 + 2]), S);
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
          kt_project(KT, mk[3 * ll], mk[3 * ll + 1], mk[3 * ll + 2], bu[m], bv[m]);
          ++m;
        }
        // nearest back-projection for every unused detection (:664, calculateMinDistancesAndPairs :862-906)
        uint32_t within = 0;            // bit i: unused detection i is within tolerance
        unsigned long long pairs = 0;   // 4 bits per unused detection: index of the nearest unused LED
        int ui = 0;
        for (int kk = 0; kk < n_det; ++kk) {
          if (kk == d0 || kk == d1 || kk == d2) continue;
          const double du = det[2 * kk], dv = det[2 * kk + 1];
          double best = HUGE_VAL;
          int bj = 0;
          for (int j = 0; j < n_unused_obj; ++j) {
            double dx = du - bu[j], dy = dv - bv[j];
            double d2v = dx * dx + dy * dy;
            if (d2v < best) { best = d2v; bj = j; }
          }
          if (best <= tol_sq_max) within |= 1u << ui;        // :671  sqrt(best) < tol, see DevPoseParams::back_proj_sq_max
          pairs |= (unsigned long long)bj << (4 * ui);
          ++ui;
        }
        if (within) {                                         // :676
          atomicAdd(&ghist[d0 * n_obj + o0], 1u);             // :680-685
          atomicAdd(&ghist[d1 * n_obj + o1], 1u);
          atomicAdd(&ghist[d2 * n_obj + o2], 1u);
          ui = 0;
          for (int kk = 0; kk < n_det; ++kk) {                // :687-695
            if (kk == d0 || kk == d1 || kk == d2) continue;
            if (within & (1u << ui)) {
              int bj = (int)((pairs >> (4 * ui)) & 0xf);
              int obj = -1, cnt = 0;                          // bj-th unused LED -> LED index
              for (int ll = 0; ll < n_obj; ++ll) {
                if (ll == o0 || ll == o1 || ll == o2) continue;
                if (cnt == bj) { obj = ll; break; }
                ++cnt;
              }
              atomicAdd(&ghist[kk * n_obj + obj], 1u);
            }
            ++ui;
          }
        }
      }
    }
    rot = (rot + count) % kK2Threads;
    __syncwarp();                                       // reconverge the warp before the next frame (lanes ran 2 or 3 problems)
  }
}

// correspondencesFromHistogram for every frame (one thread each), after the sweep
__global__ void decode_kernel(const K2Args a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n_frames) return;
  if (a.active && !a.active[f]) return;
  const int n_det = a.n_det[f], n_obj = a.pp.n_obj;
  if (n_det < 4 || n_det > MPE_MAX_DET) {
    a.n_corr[f] = 0;
    if (n_det > MPE_MAX_DET) a.frame_flags[f] |= MPE_F_TOO_MANY_DET;
    return;
  }
  const uint32_t* ghist = a.hist + (size_t)f * MPE_MAX_DET * MPE_MAX_LEDS;
  uint32_t h[MPE_MAX_DET * MPE_MAX_LEDS];
  bool all_zero = true;                                       // pose_estimator.cpp:704: decode only if the histogram is not all zero
  for (int i = 0; i < n_det * n_obj; ++i) { h[i] = ghist[i]; all_zero = all_zero && (h[i] == 0); }
  int n = 0;
  if (!all_zero) n = decode_histogram(h, n_det, n_obj, a.pp.histogram_threshold, a.corr + (size_t)f * 2 * MPE_MAX_LEDS);
  a.n_corr[f] = n;
  a.frame_flags[f] |= MPE_F_INITIALISED;
}

cudaError_t launch_p3p_sweep(const K2Args& a, int n_sms, cudaStream_t st) {
  bearings_kernel<<<(a.n_frames * MPE_MAX_DET + 255) / 256, 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n_units = a.n_frames * a.split;
  static int mult = -1;
  if (mult < 0) { const char* e = getenv("MPE_K2_GRID_MULT"); mult = e ? atoi(e) : 0; }
  // default: one CTA per (frame, part) — measured faster than fewer persistent CTAs (hardware CTA scheduling balances the
  // tail; 8192 frames: 2.07 ms against 2.31 ms with 4 CTAs per resident slot); MPE_K2_GRID_MULT overrides for experiments
  int grid = (mult == 0) ? n_units : n_sms * MPE_K2_MINBLOCKS * mult;
  if (grid > n_units) grid = n_units;
  p3p_sweep_kernel<<<grid, kK2Threads, 0, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  decode_kernel<<<(a.n_frames + 127) / 128, 128, 0, st>>>(a);
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
