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
#include "tracking_math.cuh"

namespace mpe {

namespace {
}  // namespace

// ---- step 1: predictWithROI (or the cold-branch preamble) ---------------------------------------------------------
__global__ void track_begin_kernel(const TrackArgs a) {
  pdl_enter();
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
      pred = predict_pose(prev, cur, st.previous_time, st.current_time, st.predicted_time);
      for (int i = 0; i < 16; ++i) st.predicted_pose[i] = pred.m[i];
    } else {
      st.predicted_time = t;
    }
    // predictMarkerPositionsInImage (:270-276) + determineROI (led_detector.cpp:114-179)
    double* pp = a.pred_px + (size_t)s * MPE_MAX_LEDS * 2;
    for (int i = 0; i < a.pp.n_obj; ++i)
      project2d(a.cam.K, pred, a.pp.markers[3 * i], a.pp.markers[3 * i + 1], a.pp.markers[3 * i + 2], pp[2 * i], pp[2 * i + 1]);
    roi = determine_roi(a.cam, pp, a.pp.n_obj, a.img_w, a.img_h, a.roi_border);
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
  if (n > MPE_MAX_DET && n <= MPE_MAX_BLOBS) {
    // Capacity path (flagged MPE_F_TOO_MANY_DET): the brute-force tables hold 16 detections, findCorrespondences does not need
    // them.  The nearest neighbours are searched among ALL detections as the reference does (:372-392); the matched ones are
    // then compacted to the front of the stream's detection list (ascending, so in place), and checkCorrespondences /
    // optimisePose — and initialise() if the check fails — continue on that list.  Correspondence rows and the record's
    // detections refer to the compacted list.
    double* det = const_cast<double*>(a.det) + (size_t)s * MPE_MAX_BLOBS * 2;
    float* cen = const_cast<float*>(a.centers) + (size_t)s * MPE_MAX_BLOBS * 2;
    uint32_t* corr = a.corr + (size_t)s * 2 * MPE_MAX_LEDS;
    unsigned long long used = 0;                                     // bit j: detection j is some LED's nearest neighbour within tolerance
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
      if (sqrt(best) <= a.pp.nearest_neighbour_pixel_tolerance) { corr[2 * k] = (uint32_t)i + 1; corr[2 * k + 1] = bj; used |= 1ull << (bj - 1); ++k; }
    }
    int m = 0;
    for (int j = 0; j < n; ++j) {
      if (!((used >> j) & 1ull)) continue;
      det[2 * m] = det[2 * j]; det[2 * m + 1] = det[2 * j + 1]; cen[2 * m] = cen[2 * j]; cen[2 * m + 1] = cen[2 * j + 1];
      ++m;
    }
    for (int r = 0; r < k; ++r) corr[2 * r + 1] = (uint32_t)__popcll(used & ((1ull << (corr[2 * r + 1] - 1)) - 1ull)) + 1u;
    const_cast<int*>(a.n_det)[s] = m;
    a.n_corr[s] = k;
    if (m >= 4) a.a_check[s] = 1; else a.done[s] = 1;
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
  pdl_enter();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  track_after_detect_body(a, pass, s);
}

// pass 0 wrapper: afterwards the ROI of every stream that does not retry is emptied, so that the retry launches of K1 find no
// tile to work on for it
__global__ void track_after_detect0_kernel(const TrackArgs a) {
  pdl_enter();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n) return;
  track_after_detect_body(a, 0, s);
  if (!a.a_retry[s]) { Roi none; none.x = 0; none.y = 0; none.w = 0; none.h = 0; a.rois[s] = none; }
}

// (steps 3 and 4 — "checkCorrespondences succeeded -> optimisePose, else initialise()" (:835-846) and "initialise() succeeded ->
//  optimisePose" — are written by the check kernel's epilogue: K3Args::set_gn_if_ok / set_init_if_fail.)

// ---- step 5: optimiseAndUpdatePose bookkeeping (:802-812, :794-800) + result records ------------------------------
__device__ __forceinline__ void track_finish_body(const TrackArgs& a, mpe_result* out, int s) {
  StreamState& st = a.state[s];
  // short step: this stream needs a stage the short step does not contain (cold start, whole-image retry, re-initialisation) -> state
  // untouched, the host repeats the step in full (track_begin recomputes the same prediction from the unchanged poses)
  const bool needs_full = a.fast && (a.mode[s] == 0 || a.a_retry[s] || a.a_init[s]);
  const bool upd = !needs_full && a.a_gn[s] && a.updated[s];
  if (needs_full) {
  } else if (upd) {
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
  r.flags = a.flags[s] | a.track_flags[s] | (a.a_init[s] ? MPE_F_INITIALISED : 0) | (needs_full ? kFlagNeedsFullStep : 0);
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

// out_host (optional): a second copy of the records, written by the whole CTA straight into page-locked host memory — for a few
// cameras that replaces the device-to-host copy that would follow.
__global__ void track_finish_kernel(const TrackArgs a, mpe_result* out, mpe_result* out_host) {
  pdl_enter();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < a.n) track_finish_body(a, out, s);
  if (out_host) {
    static_assert(sizeof(mpe_result) % 8 == 0, "record copied in 8-byte words");
    __syncthreads();
    const int s0 = blockIdx.x * blockDim.x;
    const int cnt = min((int)blockDim.x, a.n - s0);
    const size_t words = (size_t)max(cnt, 0) * (sizeof(mpe_result) / 8);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(out + s0);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out_host + s0);
    for (size_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
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

cudaError_t launch_track_begin(const TrackArgs& a, cudaStream_t st) { return launch_k(track_begin_kernel, grid_for(a.n), 128, 0, st, a); }
cudaError_t launch_track_after_detect(const TrackArgs& a, int pass, cudaStream_t st) {
  if (pass == 0) return launch_k(track_after_detect0_kernel, grid_for(a.n), 128, 0, st, a);
  return launch_k(track_after_detect_kernel, grid_for(a.n), 128, 0, st, a, pass);
}
cudaError_t launch_track_finish(const TrackArgs& a, mpe_result* out, mpe_result* out_host, cudaStream_t st) {
  return launch_k(track_finish_kernel, grid_for(a.n), 128, 0, st, a, out, out_host);
}
cudaError_t launch_track_reset(StreamState* s, int n, cudaStream_t st) { track_reset_kernel<<<grid_for(n), 128, 0, st>>>(s, n); return cudaGetLastError(); }

}  // namespace mpe
