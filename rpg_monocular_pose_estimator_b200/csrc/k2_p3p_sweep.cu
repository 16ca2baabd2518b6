// K2 — PoseEstimator::initialise (reference: monocular_pose_estimator_lib/src/pose_estimator.cpp:544-721):
// brute-force correspondence search.  For every 3-subset of the detections (lexicographic,
// Combinations::combinationsNoReplacement) and every ordered 3-tuple of LEDs
// (Combinations::permutationsNoReplacement) one Kneip P3P problem is solved, each of its four solutions is
// scored by back-projecting the unused LEDs and voting into the n_det x n_obj integer histogram.  The
// histogram is then decoded by PoseEstimator::correspondencesFromHistogram (:344-370).
//
// Mapping: one THREAD per P3P problem (the whole solve + the four scorings are straight-line scalar FP64
// code, so a warp per problem would idle 31 lanes; with thousands of frames per launch there is no shortage
// of parallelism).  Votes are integer atomics on the frame's global histogram (order-independent, hence exact).
// What does not depend on the pairing is hoisted: the LED-triple geometry into a table built by mpe_set_markers, the
// detection-triple geometry into a per-frame table written by a prologue kernel; hypotheses pass a conservative reject
// filter and only the survivors are scored with the reference's exact arithmetic.  The histogram is decoded by a
// one-thread-per-frame epilogue kernel.
//
// FP64 / latency bound, not HBM bound: it reads < 1 KB per frame.  Compiled with -fmad=false.
#include "mpe_internal.cuh"
#include "p3p_device.cuh"
#include "p3p_tier1.cuh"
#include <cstdlib>

namespace mpe {

#ifndef MPE_K2_ROLL_K
#define MPE_K2_ROLL_K 0     // 1: keep the four back-substitutions rolled (smaller code); measured slower
#endif
#ifndef MPE_K2_THREADS
#define MPE_K2_THREADS 256
#endif
#ifndef MPE_K2_MINBLOCKS
#define MPE_K2_MINBLOCKS 3
#endif
#ifndef MPE_T1_UNROLL_K
#define MPE_T1_UNROLL_K 0   // 1: the four roots of tier 1 unrolled (more registers; measured slower at 80 registers)
#endif
#ifndef MPE_T1_FP32
#define MPE_T1_FP32 0       // 1: tier 1 projects the unused LEDs and compares in single precision (p3p_tier1.cuh; validated, 4 % faster, 3x less headroom under the margin); 0 (default): all double
#endif
#ifndef MPE_T1_DET_REGS
#define MPE_T1_DET_REGS 1   // 1: unused detections in registers when n_det == n_obj
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
// block j/6 = lexicographic combination (a<b<c); rows [c b a],[c a b],[b c a],[b a c],[a b c],[a c b].
// Branch-free (lanes of a warp hold different rows): the row pattern is a packed table of positions in (a,b,c).
__device__ __forceinline__ int pick3(int i, int a, int b, int c) { return i == 0 ? a : (i == 1 ? b : c); }
__device__ __forceinline__ void unrank_perm3(int n, int j, int& p0, int& p1, int& p2, int& a, int& b, int& c) {
  unrank_comb3(n, j / 6, a, b, c);
  const int r6 = j % 6;
  // six rows x three positions x two bits (p0 lowest): (2,1,0) (2,0,1) (1,2,0) (1,0,2) (0,1,2) (0,2,1) -> 0x06 0x12 0x09 0x21 0x24 0x18
  const uint32_t e = (uint32_t)(0x624849486ull >> (6 * r6)) & 63u;
  p0 = pick3(e & 3, a, b, c); p1 = pick3((e >> 2) & 3, a, b, c); p2 = pick3((e >> 4) & 3, a, b, c);
}
__device__ __forceinline__ void unrank_perm3(int n, int j, int& p0, int& p1, int& p2) {
  int a, b, c;
  unrank_perm3(n, j, p0, p1, p2, a, b, c);
}

// m-th (0-based) index of {0,1,2,...} that is none of a < b < c — the same for every lane's trip count, so loops over the
// unused LEDs / detections stay convergent although each lane excludes a different triple
__device__ __forceinline__ int nth_unused(int m, int a, int b, int c) {
  m += (m >= a);
  m += (m >= b);
  m += (m >= c);
  return m;
}

// PoseEstimator::calculateImageVectors (pose_estimator.cpp:288-301)
__device__ __forceinline__ v3 bearing_vector(const DevCamera& cam, double u, double v) {
  double x = (u - cam.K[2]) / cam.K[0];
  double y = (v - cam.K[5]) / cam.K[4];
  double z = 1;
  double n = sqrt(x * x + y * y + z * z);
  return v_make(x / n, y / n, z / n);
}

// ---- hoisted tables -------------------------------------------------------------------------------------------------
// (1) LED-triple table, rebuilt by mpe_set_markers: for every row j of permutationsNoReplacement(n_obj,3) the part of
//     computePoses that depends on the ordered world-point triple only (P3PWorld).  SoA: field-major [kTripleFields][n_perm],
//     so that lanes with consecutive j load consecutive words.  Field 15 = code: 0 colinear (p3p.cpp:77-80), 1 usable,
//     2 usable and well conditioned (the conservative reject filter may be applied).
// (2) detection-triple table, rebuilt every frame by the prologue: for every 3-subset of the detections the bearings-only
//     part (P3PCamera).  AoS [frame][combo][kComboFields]: the lanes of a warp mostly share the combo (broadcast loads).
//     Field 12 = code: bit 0 swap (p3p.cpp:101-121), bit 1 well conditioned.
constexpr double kCondMin = 1e-6;        // sine of the angle that defines a frame; below it the filter is not trusted

__global__ void marker_triples_kernel(const DevPoseParams pp, int n_perm, double* __restrict__ tab) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_perm) return;
  int o0, o1, o2, oa, ob, oc;
  unrank_perm3(pp.n_obj, j, o0, o1, o2, oa, ob, oc);
  const double* mk = pp.markers;
  P3PWorld W;
  p3p_world_frame(v_make(mk[3 * o0], mk[3 * o0 + 1], mk[3 * o0 + 2]), v_make(mk[3 * o1], mk[3 * o1 + 1], mk[3 * o1 + 2]),
                  v_make(mk[3 * o2], mk[3 * o2 + 1], mk[3 * o2 + 2]), W);
  const double f[15] = {W.n1.x, W.n1.y, W.n1.z, W.n2.x, W.n2.y, W.n2.z, W.n3.x, W.n3.y, W.n3.z, W.P1.x, W.P1.y, W.P1.z,
                        W.p_1, W.p_2, W.d_12};
  for (int k = 0; k < 15; ++k) tab[(size_t)k * n_perm + j] = f[k];
  double code = 0.0;
  if (!(W.cross_norm == 0.0)) {
    const double len13 = sqrt(W.p_1 * W.p_1 + W.p_2 * W.p_2);        // |P3 - P1| (P3 has no z component in the world frame)
    code = (W.cross_norm > kCondMin * W.d_12 * len13) ? 2.0 : 1.0;
  }
  tab[(size_t)15 * n_perm + j] = code;
  tab[(size_t)16 * n_perm + j] = (double)(o0 | (o1 << 4) | (o2 << 8) | (oa << 12) | (ob << 16) | (oc << 20));
  // tier 1: the unused LEDs in this triple's world frame, X_N = N (X - P1), in LED order
  const int nu = pp.n_obj - 3;
  for (int m = 0; m < nu; ++m) {
    const int ll = nth_unused(m, oa, ob, oc);
    const v3 dx = v_sub(v_make(mk[3 * ll], mk[3 * ll + 1], mk[3 * ll + 2]), W.P1);
    tab[(size_t)(kTripleXN + 3 * m) * n_perm + j] = v_dot(W.n1, dx);
    tab[(size_t)(kTripleXN + 3 * m + 1) * n_perm + j] = v_dot(W.n2, dx);
    tab[(size_t)(kTripleXN + 3 * m + 2) * n_perm + j] = v_dot(W.n3, dx);
  }
}

cudaError_t launch_marker_triples(const DevPoseParams& pp, double* table, cudaStream_t st) {
  const int n = pp.n_obj;
  const int n_perm = n * (n - 1) * (n - 2);
  if (n_perm <= 0) return cudaSuccess;
  marker_triples_kernel<<<(n_perm + 127) / 128, 128, 0, st>>>(pp, n_perm, table);
  return cudaGetLastError();
}

// setImagePoints -> image_vectors_ (pose_estimator.cpp:166-170, 288-301) and the bearings-only part of every P3P problem of
// the frame: kComboLanes threads per frame, each walking over the frame's 3-subsets of the detections
constexpr int kComboLanes = 32;
__global__ void combo_setup_kernel(const K2Args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = i / kComboLanes, lane = i - f * kComboLanes;
  if (f >= a.n_frames) return;
  if (a.active && !a.active[f]) return;
  const int n_det = a.n_det[f];
  if (n_det < 4 || n_det > MPE_MAX_DET) return;
  const int n_comb = n_det * (n_det - 1) * (n_det - 2) / 6;
  const double* det = a.det + (size_t)f * a.det_stride * 2;
  for (int ci = lane; ci < n_comb; ci += kComboLanes) {
    int d0, d1, d2;
    unrank_comb3(n_det, ci, d0, d1, d2);
    P3PCamera Cm;
    p3p_camera_frame(bearing_vector(a.cam, det[2 * d0], det[2 * d0 + 1]), bearing_vector(a.cam, det[2 * d1], det[2 * d1 + 1]),
                     bearing_vector(a.cam, det[2 * d2], det[2 * d2 + 1]), Cm);
    double* o = a.combos + ((size_t)f * kMaxCombos + ci) * kComboFields;
    o[0] = Cm.e1.x; o[1] = Cm.e1.y; o[2] = Cm.e1.z; o[3] = Cm.e2.x; o[4] = Cm.e2.y; o[5] = Cm.e2.z;
    o[6] = Cm.e3.x; o[7] = Cm.e3.y; o[8] = Cm.e3.z; o[9] = Cm.f_1; o[10] = Cm.f_2; o[11] = Cm.b;
    o[12] = (double)(Cm.swap | ((Cm.sin12 > kCondMin) ? 2 : 0));
    o[13] = (double)(d0 | (d1 << 4) | (d2 << 8));
    // tier 1: Mc = K [e1 e2 e3] (columns), row-major
    const double* K = a.cam.K;
    const v3 e[3] = {Cm.e1, Cm.e2, Cm.e3};
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) o[14 + 3 * r + q] = K[3 * r] * e[q].x + K[3 * r + 1] * e[q].y + K[3 * r + 2] * e[q].z;
  }
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

// frames that take part in this sweep (active and with enough detections), in any order
__global__ void compact_active_kernel(const K2Args a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n_frames) return;
  if (!a.active[f]) return;
  const int n_det = a.n_det[f];
  if (n_det < 4 || n_det > MPE_MAX_DET) return;
  a.frame_list[atomicAdd(a.frame_count, 1u)] = f;
}

// Exact scoring of one finite pose hypothesis H = [R|C] (camera -> world), exactly as PoseEstimator::initialise does it
// (pose_estimator.cpp:656-697): general inverse, project2d of the unused LEDs, nearest back-projection for every unused
// detection, votes.  ids packs d0,d1,d2,o0,o1,o2 (4 bits each).  Not inlined: it runs for ~1 % of the hypotheses.
__device__ __noinline__ void score_and_vote(const double H[12], uint32_t ids, int n_det, int n_obj, const double* __restrict__ det,
                                            const double* __restrict__ mk, const double* __restrict__ K, double tol_sq_max,
                                            uint32_t* __restrict__ ghist) {
  const int d0 = ids & 15, d1 = (ids >> 4) & 15, d2 = (ids >> 8) & 15;
  const int o0 = (ids >> 12) & 15, o1 = (ids >> 16) & 15, o2 = (ids >> 20) & 15;
  const int n_unused_obj = n_obj - 3;
  double Hi[12], KT[12];
  h_inverse(H, Hi);                                    // :660
  kt_product(K, Hi, KT);
  // sorted copy of the LED triple (the exclusion list of the loops below)
  int oa = min(o0, min(o1, o2)), oc = max(o0, max(o1, o2)), ob = o0 + o1 + o2 - oa - oc;
  // back-project the unused LEDs (:658-661), in LED order
  double bu[MPE_MAX_LEDS - 3], bv[MPE_MAX_LEDS - 3];
  for (int m = 0; m < n_unused_obj; ++m) {
    const int ll = nth_unused(m, oa, ob, oc);
    kt_project(KT, mk[3 * ll], mk[3 * ll + 1], mk[3 * ll + 2], bu[m], bv[m]);
  }
  // nearest back-projection for every unused detection (:664, calculateMinDistancesAndPairs :862-906)
  uint32_t within = 0;            // bit i: unused detection i is within tolerance
  unsigned long long pairs = 0;   // 4 bits per unused detection: index of the nearest unused LED
  const int n_unused_det = n_det - 3;
  for (int ui = 0; ui < n_unused_det; ++ui) {
    const int kk = nth_unused(ui, d0, d1, d2);
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
  }
  if (within) {                                         // :676
    atomicAdd(&ghist[d0 * n_obj + o0], 1u);             // :680-685
    atomicAdd(&ghist[d1 * n_obj + o1], 1u);
    atomicAdd(&ghist[d2 * n_obj + o2], 1u);
    for (int ui = 0; ui < n_unused_det; ++ui) {         // :687-695
      if (within & (1u << ui)) {
        const int kk = nth_unused(ui, d0, d1, d2);
        const int obj = nth_unused((int)((pairs >> (4 * ui)) & 0xf), oa, ob, oc);   // bj-th unused LED -> LED index
        atomicAdd(&ghist[kk * n_obj + obj], 1u);
      }
    }
  }
}

// Conservative reject test: false only if NO unused detection can lie within the back-projection tolerance of ANY unused
// LED under this hypothesis, so that skipping the exact evaluation cannot change a vote.  It projects with the rigid inverse
// R^T (X - C) instead of the general 4x4 inverse and compares without dividing:  |K x_c - (u,v,1) z|^2 <= r^2 z^2  with
// r = tolerance + margin.  H is a product of three frames that are orthonormal to ~1e-10 when both conditioning codes are
// set, so the two projections differ by < 4e-6 * fx pixels for |z| >= 1e-3 |x_c|_1 (closer to the camera plane, or any NaN,
// the test answers "maybe"); the host disables the filter unless that is far below the margin.
// `bb` = centre and half extent of the bounding box of ALL detections of the frame (uc, vc, hu, hv): with many unused detections
// a projection outside the box grown by r skips its pair loop (kBBox: compiled in only for objects with >= 7 LEDs, where the pair
// loop is long; for few LEDs the box would only cost registers).
template <bool kBBox>
__device__ __forceinline__ bool maybe_within(const double H[12], int oa, int ob, int oc, int d0, int d1, int d2, int n_det, int n_obj,
                                             const double* __restrict__ det, const double* __restrict__ mk, const double* __restrict__ K,
                                             double r, const double bb[4]) {
  bool maybe = false;
  const int nu_obj = n_obj - 3, nu_det = n_det - 3;
  const double r2 = r * r;
  for (int m = 0; m < nu_obj; ++m) {
    const int ll = nth_unused(m, oa, ob, oc);            // oa < ob < oc: the LED triple of this problem, sorted
    const double dx = mk[3 * ll] - H[3], dy = mk[3 * ll + 1] - H[7], dz = mk[3 * ll + 2] - H[11];
    const double xc = H[0] * dx + H[4] * dy + H[8] * dz;
    const double yc = H[1] * dx + H[5] * dy + H[9] * dz;
    const double zc = H[2] * dx + H[6] * dy + H[10] * dz;
    const double au = K[0] * xc + K[1] * yc + K[2] * zc;
    const double av = K[3] * xc + K[4] * yc + K[5] * zc;
    const double az = K[6] * xc + K[7] * yc + K[8] * zc;
    const bool near_plane = !(fabs(az) >= 1e-3 * (fabs(xc) + fabs(yc) + fabs(zc)));
    maybe = maybe || near_plane;
    if (kBBox) {
      const double aaz = fabs(az);
      const bool outside = fabs(au - bb[0] * az) > (bb[2] + r) * aaz || fabs(av - bb[1] * az) > (bb[3] + r) * aaz;
      if (outside && !near_plane) continue;              // cannot be within r of any detection
    }
    const double lim = r2 * (az * az);
    for (int i = 0; i < nu_det; ++i) {
      const int kk = nth_unused(i, d0, d1, d2);          // d0 < d1 < d2 (a combination)
      const double eu = au - det[2 * kk] * az, ev = av - det[2 * kk + 1] * az;
      maybe = maybe || !(eu * eu + ev * ev > lim);
    }
  }
  return maybe;
}

// The sweep.  A CTA walks over the units (frame, part) blockIdx.x, blockIdx.x + gridDim.x, ... (by default exactly one) and
// treats the unit's P3P problems as one flat sequence dealt round-robin to its threads (`rot` carries the position over from
// unit to unit).  Per problem: two table look-ups (detection triple, LED triple), the quartic, four back-substitutions; each
// finite hypothesis goes through the reject filter, and the few survivors are parked in a shared-memory queue that the CTA
// drains with the exact scoring after a barrier (every kK2DrainEvery passes and at the end of the unit) — so the expensive,
// rarely needed path runs compacted instead of diverging in most warps.  Votes are integer atomics on the frame's global
// histogram (order independent).  The block size is chosen at launch so that the unit's last pass is well filled.
constexpr int kK2DrainEvery = 4;

template <bool kBBox>
__global__ void __launch_bounds__(kK2Threads, MPE_K2_MINBLOCKS) p3p_sweep_kernel(const K2Args a) {
  const int tid = threadIdx.x;
  const int n_thr = blockDim.x;
  const int n_obj = a.pp.n_obj;
  const int n_perm = n_obj * (n_obj - 1) * (n_obj - 2);
  const double tol_sq_max = a.pp.back_proj_sq_max;
  // marker coordinates and K in shared memory: lanes of a warp index the markers with different permutations, which the
  // constant bank (kernel parameters) would serialise
  __shared__ double mk[3 * MPE_MAX_LEDS];
  __shared__ double Ks[9];
  __shared__ double q_h[12][kK2Queue];                  // parked hypotheses: H, field-major
  __shared__ uint32_t q_ids[kK2Queue];
  __shared__ int q_n[2];
  if (tid < 3 * MPE_MAX_LEDS) mk[tid] = a.pp.markers[tid];
  if (tid < 9) Ks[tid] = a.cam.K[tid];
  if (tid < 2) q_n[tid] = 0;
  __syncthreads();
  const double* __restrict__ tt = a.triples;
  // units = (frame, part).  With a compacted frame list the number of parts adapts to the length of the list, so that a
  // handful of re-initialising streams still spreads over the grid.
  const bool listed = a.frame_list != nullptr;
  const int n_listed = listed ? (int)*a.frame_count : a.n_frames;
  int split = a.split;
  if (listed && n_listed > 0) { const int s2 = (int)gridDim.x / n_listed; split = max(split, min(16, s2)); }
  const int n_units = n_listed * split;
  int rot = 0;                                          // flat position (mod blockDim) where this unit's first problem falls
  int par = 0;                                          // which of the two queue counters is being filled
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int fi = unit / split, part = unit - fi * split;
    const int f = listed ? a.frame_list[fi] : fi;
    if (!listed && a.active && !a.active[f]) continue;
    const int n_det = a.n_det[f];
    if (n_det < 4 || n_det > MPE_MAX_DET) continue;     // pose_estimator.cpp:80 (min_num_leds_detected_ = 4); flagged by decode_kernel
    const int n_comb = n_det * (n_det - 1) * (n_det - 2) / 6;
    const int total = n_comb * n_perm;
    const int chunk = (total + split - 1) / split;
    const int t_begin = part * chunk, t_end = min(total, t_begin + chunk);
    const int count = max(t_end - t_begin, 0);
    const double* det = a.det + (size_t)f * a.det_stride * 2;
    const double* combos = a.combos + (size_t)f * kMaxCombos * kComboFields;
    uint32_t* ghist = a.hist + (size_t)f * MPE_MAX_DET * MPE_MAX_LEDS;
    double bb[4] = {0, 0, 0, 0};
    if (kBBox) {                                        // bounding box of the detections
      double u0 = det[0], u1 = det[0], v0 = det[1], v1 = det[1];
      for (int i = 1; i < n_det; ++i) { u0 = fmin(u0, det[2 * i]); u1 = fmax(u1, det[2 * i]); v0 = fmin(v0, det[2 * i + 1]); v1 = fmax(v1, det[2 * i + 1]); }
      bb[0] = 0.5 * (u0 + u1); bb[1] = 0.5 * (v0 + v1);
      bb[2] = 0.5 * (u1 - u0) * (1.0 + 1e-12) + 1e-9; bb[3] = 0.5 * (v1 - v0) * (1.0 + 1e-12) + 1e-9;   // rounding of centre / half extent
    }

    int first = tid - rot;
    if (first < 0) first += n_thr;
    const int n_pass = (count + n_thr - 1) / n_thr;       // uniform over the CTA; a thread's problems are first, first + n_thr, ...
    for (int pass = 0; pass < n_pass; ++pass) {
      const int t = t_begin + first + pass * n_thr;
      if (t < t_end) {
        const int ci = t / n_perm, pj = t - ci * n_perm;
        if (tt[(size_t)15 * n_perm + pj] != 0.0) {          // else: colinear LED triple (tested on the unswapped order, p3p.cpp:77-80)
          const double* cb = combos + (size_t)ci * kComboFields;
          const int ccode = (int)cb[12];
          // the exchange of points 1 and 2 selects the table row of the exchanged triple: within a block of six rows
          // [c b a],[c a b],[b c a],[b a c],[a b c],[a c b] the partner rows are 0<->2, 1<->5, 3<->4
          int tj = pj;
          if (ccode & 1) { const int r6 = pj % 6; tj = pj - r6 + ((0x134052 >> (4 * r6)) & 7); }
          const bool filt = a.use_filter && (ccode & 2) && tt[(size_t)15 * n_perm + tj] == 2.0;
          P3PSetup S;
          // the six scalars the quartic needs first; the 21 frame entries only after it (they would just be spilled across the
          // ~1000 instructions of the quartic at 80 registers per thread — the barrier keeps the compiler from hoisting the loads)
          S.f_1 = cb[9]; S.f_2 = cb[10]; S.b = cb[11];
          S.p_1 = tt[(size_t)12 * n_perm + tj]; S.p_2 = tt[(size_t)13 * n_perm + tj]; S.d_12 = tt[(size_t)14 * n_perm + tj];
          p3p_quartic(S.f_1, S.f_2, S.p_1, S.p_2, S.d_12, S.b, S.roots);
          asm volatile("" ::: "memory");
          S.e1 = v_make(cb[0], cb[1], cb[2]); S.e2 = v_make(cb[3], cb[4], cb[5]); S.e3 = v_make(cb[6], cb[7], cb[8]);
          S.n1 = v_make(tt[tj], tt[(size_t)n_perm + tj], tt[(size_t)2 * n_perm + tj]);
          S.n2 = v_make(tt[(size_t)3 * n_perm + tj], tt[(size_t)4 * n_perm + tj], tt[(size_t)5 * n_perm + tj]);
          S.n3 = v_make(tt[(size_t)6 * n_perm + tj], tt[(size_t)7 * n_perm + tj], tt[(size_t)8 * n_perm + tj]);
          S.P1 = v_make(tt[(size_t)9 * n_perm + tj], tt[(size_t)10 * n_perm + tj], tt[(size_t)11 * n_perm + tj]);

          int d0, d1, d2, o0, o1, o2, oa, ob, oc;
          unrank_comb3(n_det, ci, d0, d1, d2);
          unrank_perm3(n_obj, pj, o0, o1, o2, oa, ob, oc);
          const uint32_t ids = (uint32_t)d0 | ((uint32_t)d1 << 4) | ((uint32_t)d2 << 8) | ((uint32_t)o0 << 12) | ((uint32_t)o1 << 16) | ((uint32_t)o2 << 20);

#if MPE_K2_ROLL_K
#pragma unroll 1
#endif
          for (int k = 0; k < 4; ++k) {
            double H[12];
            if (!p3p_solution(S, k, H)) continue;
            if (!h_is_finite(H)) continue;                       // pose_estimator.cpp:653
            if (filt && !maybe_within<kBBox>(H, oa, ob, oc, d0, d1, d2, n_det, n_obj, det, mk, Ks, a.filter_r, bb)) continue;
            const int slot = atomicAdd(&q_n[par], 1);
            if (slot < kK2Queue) {
#pragma unroll
              for (int e = 0; e < 12; ++e) q_h[e][slot] = H[e];
              q_ids[slot] = ids;
            } else {                                             // queue full: score in place (rare; the copy keeps H itself in registers)
              double Hc[12];
#pragma unroll
              for (int e = 0; e < 12; ++e) Hc[e] = H[e];
              score_and_vote(Hc, ids, n_det, n_obj, det, mk, Ks, tol_sq_max, ghist);
            }
          }
        }
      }
      const bool last = (pass + 1 == n_pass);
      if (last || (pass % kK2DrainEvery) == kK2DrainEvery - 1) {
        __syncthreads();                                  // every survivor so far is parked
        const int n_q = min(q_n[par], kK2Queue);
        if (tid == 0) q_n[par ^ 1] = 0;
        for (int e = tid; e < n_q; e += n_thr) {
          double H[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) H[i] = q_h[i][e];
          score_and_vote(H, q_ids[e], n_det, n_obj, det, mk, Ks, tol_sq_max, ghist);
        }
        par ^= 1;
        if (!last || unit + (int)gridDim.x < n_units) __syncthreads();   // the queue storage is reused
      }
    }
    rot = (rot + count) % n_thr;
  }
}

// ---- the two-tier sweep --------------------------------------------------------------------------------------------------
// Tier 1 (p3p_tier1.cuh) answers "certainly no vote" for ~95 % of the problems at about a quarter of the instructions of the
// exact solve; the rest is parked in a shared-memory queue of (frame, problem) pairs, and whenever a full block of them has
// accumulated — across frames: the CTA is persistent — every thread takes one and runs the reference's arithmetic
// (p3p_quartic, four back-substitutions, exact scoring).  A CTA flattens `group` consecutive frames into one problem sequence
// so that its passes are full (600 problems of one 5-LED frame fill 2.3 passes of 256 threads, 1200 of two fill 4.7).
// kNuObj = number of unused LEDs (n_obj - 3) when it is compiled in (loops unrolled), 0 = any number; kNuDet likewise for the
// unused detections; kXRegs: the unused LEDs' coordinates are held in registers (few LEDs) instead of being re-read from the
// table.  `sdet` = the frame's detections in shared memory, `dlist` = the indices of the unused detections (4 bits each).
template <int kNuObj, int kNuDet, bool kXRegs, bool kBBox>
__device__ __forceinline__ bool tier1_maybe(const double* __restrict__ cb, const double* __restrict__ tt, int n_perm, int tj,
                                            unsigned long long dlist, int nu_det_rt, int nu_obj_rt, const double2* __restrict__ sdet,
                                            double r, const double bb[4]) {
  T1Roots R;
  const double f_1 = cb[9], f_2 = cb[10], b = cb[11];
  const double* __restrict__ tr = tt + tj;                  // field k of this row: tr[k * n_perm]
  const double p_1 = tr[12 * n_perm], p_2 = tr[13 * n_perm], d_12 = tr[14 * n_perm];
  t1_quartic_roots(f_1, f_2, b, p_1, p_2, d_12, R);
  if (R.maybe) return true;
  const int nu_obj = (kNuObj > 0) ? kNuObj : nu_obj_rt;
  const int nu_det = (kNuDet > 0) ? kNuDet : nu_det_rt;
  T1Problem Q;
  t1_problem(f_1, f_2, b, p_1, p_2, d_12, Q);
#if MPE_T1_FP32
  // roots and root-dependent scalars in double, the projection of the unused LEDs and the comparisons in single precision
  const float rf = (float)r, r2 = rf * rf;
  float Mc[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) Mc[e] = (float)cb[14 + e];
  float X[kXRegs ? kNuObj : 1][3];
  if (kXRegs) {
#pragma unroll
    for (int m = 0; m < kNuObj; ++m)
#pragma unroll
      for (int q = 0; q < 3; ++q) X[m][q] = (float)tr[(kTripleXN + 3 * m + q) * n_perm];
  }
  float2 dreg[(kNuDet > 0) ? kNuDet : 1];
  if (kNuDet > 0) {
#pragma unroll
    for (int i = 0; i < kNuDet; ++i) { const double2 d = sdet[(int)((dlist >> (4 * i)) & 15ull)]; dreg[i] = make_float2((float)d.x, (float)d.y); }
  }
  const float d12f = (float)d_12;
  float bbf[4] = {0.f, 0.f, 0.f, 0.f};
  if (kBBox) { bbf[0] = (float)bb[0]; bbf[1] = (float)bb[1]; bbf[2] = (float)bb[2] + rf + 1e-3f; bbf[3] = (float)bb[3] + rf + 1e-3f; }
  bool maybe = false;
#if MPE_T1_UNROLL_K
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int k = 0; k < 4; ++k) {
    T1Pose Pd;
    const int st = t1_pose(R.rho[k], Q, Pd);
    if (st == 2) return true;
    if (st == 1) {
      const T1PoseF P = t1_pose_f(Pd);
#pragma unroll
      for (int m = 0; m < nu_obj; ++m) {
        float X0, X1, X2;
        if (kXRegs) { X0 = X[m][0]; X1 = X[m][1]; X2 = X[m][2]; }
        else { X0 = (float)tr[(kTripleXN + 3 * m) * n_perm]; X1 = (float)tr[(kTripleXN + 3 * m + 1) * n_perm]; X2 = (float)tr[(kTripleXN + 3 * m + 2) * n_perm]; }
        float au, av, az, l1;
        t1_project_f(P, Mc, X0, X1, X2, au, av, az, l1);
        // close to the camera plane (or to the camera itself): the division-free comparison is not trusted
        const bool near_plane = !(fabsf(az) >= 1e-3f * l1) || !(l1 >= 1e-3f * d12f);
        maybe = maybe || near_plane;
        if (kBBox) {
          const float aaz = fabsf(az);
          const bool outside = fabsf(au - bbf[0] * az) > bbf[2] * aaz || fabsf(av - bbf[1] * az) > bbf[3] * aaz;
          if (outside && !near_plane) continue;
        }
        const float lim = r2 * (az * az);
        if (kNuDet > 0) {
#pragma unroll
          for (int i = 0; i < kNuDet; ++i) {
            const float eu = T1_FMAF(-dreg[i].x, az, au), ev = T1_FMAF(-dreg[i].y, az, av);
            maybe = maybe || !(T1_FMAF(eu, eu, ev * ev) > lim);
          }
        } else {
          unsigned long long dl = dlist;
          for (int i = 0; i < nu_det; ++i, dl >>= 4) {
            const double2 d = sdet[(int)(dl & 15ull)];
            const float eu = T1_FMAF(-(float)d.x, az, au), ev = T1_FMAF(-(float)d.y, az, av);
            maybe = maybe || !(T1_FMAF(eu, eu, ev * ev) > lim);
          }
        }
      }
      if (maybe) return true;
    }
  }
  return false;
#else
  const double r2 = r * r;
  double Mc[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) Mc[e] = cb[14 + e];
  double X[kXRegs ? kNuObj : 1][3];
  if (kXRegs) {
#pragma unroll
    for (int m = 0; m < kNuObj; ++m)
#pragma unroll
      for (int q = 0; q < 3; ++q) X[m][q] = tr[(kTripleXN + 3 * m + q) * n_perm];
  }
  double2 dreg[(kNuDet > 0) ? kNuDet : 1];
  if (kNuDet > 0) {
#pragma unroll
    for (int i = 0; i < kNuDet; ++i) dreg[i] = sdet[(int)((dlist >> (4 * i)) & 15ull)];
  }
  bool maybe = false;
#if MPE_T1_UNROLL_K
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int k = 0; k < 4; ++k) {
    T1Pose P;
    const int st = t1_pose(R.rho[k], Q, P);
    if (st == 2) return true;
    if (st == 1) {
#pragma unroll
      for (int m = 0; m < nu_obj; ++m) {
        double X0, X1, X2;
        if (kXRegs) { X0 = X[m][0]; X1 = X[m][1]; X2 = X[m][2]; }
        else { X0 = tr[(kTripleXN + 3 * m) * n_perm]; X1 = tr[(kTripleXN + 3 * m + 1) * n_perm]; X2 = tr[(kTripleXN + 3 * m + 2) * n_perm]; }
        double au, av, az, l1;
        t1_project(P, Mc, X0, X1, X2, au, av, az, l1);
        // close to the camera plane (or to the camera itself): the division-free comparison is not trusted
        const bool near_plane = !(fabs(az) >= 1e-3 * l1) || !(l1 >= 1e-3 * d_12);
        maybe = maybe || near_plane;
        if (kBBox) {
          const double aaz = fabs(az);
          const bool outside = fabs(au - bb[0] * az) > (bb[2] + r) * aaz || fabs(av - bb[1] * az) > (bb[3] + r) * aaz;
          if (outside && !near_plane) continue;
        }
        const double lim = r2 * (az * az);
        if (kNuDet > 0) {
#pragma unroll
          for (int i = 0; i < kNuDet; ++i) {
            const double eu = T1_FMA(-dreg[i].x, az, au), ev = T1_FMA(-dreg[i].y, az, av);
            maybe = maybe || !(T1_FMA(eu, eu, ev * ev) > lim);
          }
        } else {
          unsigned long long dl = dlist;
          for (int i = 0; i < nu_det; ++i, dl >>= 4) {
            const double2 d = sdet[(int)(dl & 15ull)];
            const double eu = T1_FMA(-d.x, az, au), ev = T1_FMA(-d.y, az, av);
            maybe = maybe || !(T1_FMA(eu, eu, ev * ev) > lim);
          }
        }
      }
      if (maybe) return true;
    }
  }
  return false;
#endif
}

constexpr int kK2MaxGroup = 8;

template <bool kBBox>
__global__ void __launch_bounds__(kK2Threads, MPE_K2_MINBLOCKS) p3p_sweep_t1_kernel(const K2Args a) {
  const int tid = threadIdx.x;
  const int n_thr = blockDim.x;
  const int n_obj = a.pp.n_obj;
  const int n_perm = n_obj * (n_obj - 1) * (n_obj - 2);
  const float inv_perm = 1.0f / (float)n_perm;
  const double tol_sq_max = a.pp.back_proj_sq_max;
  __shared__ double mk[3 * MPE_MAX_LEDS];
  __shared__ double Ks[9];
  __shared__ uint2 sq[kK2Survivors];                     // parked problems: (frame, problem index)
  __shared__ int sq_n;
  __shared__ int g_frame[kK2MaxGroup], g_ndet[kK2MaxGroup], g_begin[kK2MaxGroup], g_offs[kK2MaxGroup + 1];
  __shared__ double2 g_det[kK2MaxGroup][MPE_MAX_DET];    // the detections of the unit's frames
  __shared__ double g_bb[kK2MaxGroup][4];
  if (tid < 3 * MPE_MAX_LEDS) mk[tid] = a.pp.markers[tid];
  if (tid < 9) Ks[tid] = a.cam.K[tid];
  if (tid == 0) sq_n = 0;
  __syncthreads();
  const double* __restrict__ tt = a.triples;
  const bool listed = a.frame_list != nullptr;
  const int n_listed = listed ? (int)*a.frame_count : a.n_frames;
  int split = a.split;
  const int group = (split > 1) ? 1 : a.group;
  if (listed && n_listed > 0 && group == 1) { const int s2 = (int)gridDim.x / n_listed; split = max(split, min(16, s2)); }
  const int n_units = (split > 1) ? n_listed * split : (n_listed + group - 1) / group;

  // exact solve of one parked problem (the reference's arithmetic all the way)
  auto exact = [&](int f, int t) {
    const int n_det = a.n_det[f];
    const int ci = t / n_perm, pj = t - ci * n_perm;
    const double* cb = a.combos + ((size_t)f * kMaxCombos + ci) * kComboFields;
    const int ccode = (int)cb[12];
    int tj = pj;
    if (ccode & 1) { const int r6 = pj % 6; tj = pj - r6 + ((0x134052 >> (4 * r6)) & 7); }
    P3PSetup S;
    S.f_1 = cb[9]; S.f_2 = cb[10]; S.b = cb[11];
    S.p_1 = tt[(size_t)12 * n_perm + tj]; S.p_2 = tt[(size_t)13 * n_perm + tj]; S.d_12 = tt[(size_t)14 * n_perm + tj];
    p3p_quartic(S.f_1, S.f_2, S.p_1, S.p_2, S.d_12, S.b, S.roots);
    asm volatile("" ::: "memory");
    S.e1 = v_make(cb[0], cb[1], cb[2]); S.e2 = v_make(cb[3], cb[4], cb[5]); S.e3 = v_make(cb[6], cb[7], cb[8]);
    S.n1 = v_make(tt[tj], tt[(size_t)n_perm + tj], tt[(size_t)2 * n_perm + tj]);
    S.n2 = v_make(tt[(size_t)3 * n_perm + tj], tt[(size_t)4 * n_perm + tj], tt[(size_t)5 * n_perm + tj]);
    S.n3 = v_make(tt[(size_t)6 * n_perm + tj], tt[(size_t)7 * n_perm + tj], tt[(size_t)8 * n_perm + tj]);
    S.P1 = v_make(tt[(size_t)9 * n_perm + tj], tt[(size_t)10 * n_perm + tj], tt[(size_t)11 * n_perm + tj]);
    const uint32_t ids = ((uint32_t)cb[13] & 0xfffu) | (((uint32_t)tt[(size_t)16 * n_perm + pj] & 0xfffu) << 12);   // d0 d1 d2 | o0 o1 o2 (unswapped row)
    const double* det = a.det + (size_t)f * a.det_stride * 2;
    uint32_t* ghist = a.hist + (size_t)f * MPE_MAX_DET * MPE_MAX_LEDS;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      double H[12];
      if (!p3p_solution(S, k, H)) continue;
      if (!h_is_finite(H)) continue;                       // pose_estimator.cpp:653
      score_and_vote(H, ids, n_det, n_obj, det, mk, Ks, tol_sq_max, ghist);
    }
  };
  auto drain = [&](bool all) {                              // called by every thread of the CTA
    __syncthreads();
    int n_q = min(sq_n, kK2Survivors);
    int done = 0;
    while (n_q - done >= n_thr || (all && done < n_q)) {
      const int e = done + tid;
      if (e < n_q) exact((int)sq[e].x, (int)sq[e].y);
      done += n_thr;
    }
    __syncthreads();
    if (done > 0) {                                         // keep the tail (fewer than a block) for the next round
      const int rest = max(n_q - done, 0);
      uint2 keep = make_uint2(0, 0);
      if (tid < rest) keep = sq[done + tid];
      __syncthreads();
      if (tid < rest) sq[tid] = keep;
      if (tid == 0) sq_n = rest;
      __syncthreads();
    }
  };

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    // ---- the frames (or the part of one frame) of this unit
    if (tid < kK2MaxGroup) {
      int f = -1, nd = 0, begin = 0, count = 0;
      if (tid < group) {
        const int fi = (split > 1) ? unit / split : unit * group + tid;
        if (fi < n_listed) {
          f = listed ? a.frame_list[fi] : fi;
          if (!listed && a.active && !a.active[f]) f = -1;
        }
        if (f >= 0) {
          nd = a.n_det[f];
          if (nd < 4 || nd > MPE_MAX_DET) { f = -1; nd = 0; }   // pose_estimator.cpp:80; flagged by decode_kernel
        }
        if (f >= 0) {
          const int total = nd * (nd - 1) * (nd - 2) / 6 * n_perm;
          if (split > 1) {
            const int part = unit - (unit / split) * split;
            const int chunk = (total + split - 1) / split;
            begin = part * chunk;
            count = max(min(total, begin + chunk) - begin, 0);
          } else {
            count = total;
          }
          const double* det = a.det + (size_t)f * a.det_stride * 2;
          double u0 = det[0], u1 = det[0], v0 = det[1], v1 = det[1];
          for (int i = 0; i < nd; ++i) {
            const double u = det[2 * i], v = det[2 * i + 1];
            g_det[tid][i] = make_double2(u, v);
            u0 = fmin(u0, u); u1 = fmax(u1, u); v0 = fmin(v0, v); v1 = fmax(v1, v);
          }
          g_bb[tid][0] = 0.5 * (u0 + u1); g_bb[tid][1] = 0.5 * (v0 + v1);           // bounding box of the detections
          g_bb[tid][2] = 0.5 * (u1 - u0) * (1.0 + 1e-12) + 1e-9; g_bb[tid][3] = 0.5 * (v1 - v0) * (1.0 + 1e-12) + 1e-9;
        }
      }
      g_frame[tid] = f; g_ndet[tid] = nd; g_begin[tid] = begin;
      g_offs[tid + 1] = count;
    }
    __syncthreads();
    if (tid == 0) {
      g_offs[0] = 0;
      for (int g = 0; g < kK2MaxGroup; ++g) g_offs[g + 1] += g_offs[g];
    }
    __syncthreads();
    const int total = g_offs[kK2MaxGroup];
    const int n_pass = (total + n_thr - 1) / n_thr;
    for (int pass = 0; pass < n_pass; ++pass) {
      const int q = pass * n_thr + tid;
      if (q < total) {
        int g = 0;
        if (group > 1) {
#pragma unroll
          for (int i = 1; i < kK2MaxGroup; ++i) g += (q >= g_offs[i]);
        }
        const int f = g_frame[g], n_det = g_ndet[g];
        const int t = g_begin[g] + (q - g_offs[g]);
        int ci = (int)((float)t * inv_perm);                  // t / n_perm (t < 2^21: the float quotient is off by at most one)
        int pj = t - ci * n_perm;
        if (pj < 0) { --ci; pj += n_perm; } else if (pj >= n_perm) { ++ci; pj -= n_perm; }
        const double* __restrict__ cb = a.combos + ((size_t)f * kMaxCombos + ci) * kComboFields;
        if (tt[15 * n_perm + pj] != 0.0) {                    // else: colinear LED triple (tested on the unswapped order, p3p.cpp:77-80)
          const int ccode = (int)cb[12];
          int tj = pj;
          if (ccode & 1) { const int r6 = pj % 6; tj = pj - r6 + ((0x134052 >> (4 * r6)) & 7); }
          bool survive = true;
          if ((ccode & 2) && tt[15 * n_perm + tj] == 2.0) {
            const uint32_t dd = (uint32_t)cb[13];
            const int d0 = dd & 15, d1 = (dd >> 4) & 15, d2 = (dd >> 8) & 15;
            const int nu_det = n_det - 3;
            unsigned long long dlist = 0;
            for (int i = 0; i < nu_det; ++i) dlist |= (unsigned long long)nth_unused(i, d0, d1, d2) << (4 * i);
            const double* bb = g_bb[g];
            const int nuo = n_obj - 3;
            const double r = a.filter_r;
            const double2* sd = g_det[g];
            if (nuo == 1) survive = (MPE_T1_DET_REGS && nu_det == 1) ? tier1_maybe<1, 1, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb)
                                                  : tier1_maybe<1, 0, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb);
            else if (nuo == 2) survive = (MPE_T1_DET_REGS && nu_det == 2) ? tier1_maybe<2, 2, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb)
                                                       : tier1_maybe<2, 0, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb);
            else if (nuo == 3) survive = (MPE_T1_DET_REGS && nu_det == 3) ? tier1_maybe<3, 3, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb)
                                                       : tier1_maybe<3, 0, true, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb);
            else if (nuo == 5) survive = tier1_maybe<5, 0, false, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb);
            else survive = tier1_maybe<0, 0, false, kBBox>(cb, tt, n_perm, tj, dlist, nu_det, nuo, sd, r, bb);
          }
          if (survive) {
            const int slot = atomicAdd(&sq_n, 1);
            if (slot < kK2Survivors) sq[slot] = make_uint2((uint32_t)f, (uint32_t)t);
            else exact(f, t);                                  // queue full (cannot happen with one drain per pass; kept for safety)
          }
        }
      }
      // a pass adds at most n_thr entries: drain whenever another pass might not fit
      __syncthreads();
      if (sq_n > kK2Survivors - n_thr) drain(false);
    }
    __syncthreads();                                           // g_* are rewritten by the next unit
  }
  drain(true);
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
  combo_setup_kernel<<<(int)(((size_t)a.n_frames * kComboLanes + 255) / 256), 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n_units = a.n_frames * a.split;
  static int mult = -1, forced_bs = -1;
  if (mult < 0) { const char* e = getenv("MPE_K2_GRID_MULT"); mult = e ? atoi(e) : 0; }
  if (forced_bs < 0) { const char* e = getenv("MPE_K2_THREADS"); forced_bs = e ? atoi(e) : 0; }
  // default: one CTA per (frame, part) — measured faster than fewer persistent CTAs (hardware CTA scheduling balances the
  // tail); MPE_K2_GRID_MULT overrides for experiments
  int grid = (mult == 0) ? n_units : n_sms * MPE_K2_MINBLOCKS * mult;
  if (a.frame_list) {                                     // masked sweep (tracking step): compact the frames that take part, persistent CTAs walk the list
    compact_active_kernel<<<(a.n_frames + 255) / 256, 256, 0, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    grid = n_sms * MPE_K2_MINBLOCKS * 2;
  }
  if (grid > n_units) grid = n_units;
  // block size: 256 threads (measured best for 600 and 18 816 problems per frame: the kernel is latency bound, a fuller last
  // pass with 160 threads and five CTAs per SM was 3 % slower); units with fewer problems than that get a CTA of their size
  int bs = kK2Threads;
  {
    const int n = a.pp.n_obj;
    const long long total = (long long)n * (n - 1) * (n - 2) / 6 * n * (n - 1) * (n - 2);
    const long long per_unit = (total + a.split - 1) / a.split;
    if (per_unit < kK2Threads) bs = (int)((per_unit + 31) / 32 * 32);
    if (bs < 64) bs = 64;
    if (forced_bs >= 32 && forced_bs <= kK2Threads && forced_bs % 32 == 0) bs = forced_bs;
  }
  if (a.use_filter == 2) {                                // two-tier sweep: persistent CTAs, `group` frames flattened per unit
    K2Args b = a;
    const int n = a.pp.n_obj;
    const long long per_frame = (long long)n * (n - 1) * (n - 2) / 6 * n * (n - 1) * (n - 2);
    int group = 1;
    if (a.split <= 1) {                                   // the group size that fills the passes best (ties: the smaller one)
      double best = 0;
      for (int g = 1; g <= kK2MaxGroup; ++g) {
        const long long tot = per_frame * g;
        const double fill = (double)tot / (double)((tot + kK2Threads - 1) / kK2Threads * kK2Threads);
        if (fill > best + 0.02) { best = fill; group = g; }
      }
      static int forced_group = -1;
      if (forced_group < 0) { const char* e = getenv("MPE_K2_GROUP"); forced_group = e ? atoi(e) : 0; }
      if (forced_group >= 1 && forced_group <= kK2MaxGroup) group = forced_group;
    }
    b.group = group;
    const int units = (a.split > 1) ? a.n_frames * a.split : (a.n_frames + group - 1) / group;
    // persistent CTAs: MPE_K2_MINBLOCKS per SM fill the register file (3 x 256 threads x 80 registers), so nothing of another
    // stream runs beside the sweep; MPE_K2_CTAS_PER_SM = 1 / 2 leaves room (experiments)
    static int ctas_per_sm = -1;
    if (ctas_per_sm < 0) { const char* e = getenv("MPE_K2_CTAS_PER_SM"); ctas_per_sm = e ? atoi(e) : 0; }
    int g2 = n_sms * ((ctas_per_sm >= 1 && ctas_per_sm <= MPE_K2_MINBLOCKS) ? ctas_per_sm : MPE_K2_MINBLOCKS);
    if (mult > 0) g2 *= mult;
    if (g2 > units) g2 = units;
    if (g2 < 1) g2 = 1;
    if (a.pp.n_obj >= 7) p3p_sweep_t1_kernel<true><<<g2, kK2Threads, 0, st>>>(b);
    else p3p_sweep_t1_kernel<false><<<g2, kK2Threads, 0, st>>>(b);
  } else if (a.pp.n_obj >= 7) p3p_sweep_kernel<true><<<grid, bs, 0, st>>>(a);
  else p3p_sweep_kernel<false><<<grid, bs, 0, st>>>(a);
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
