/*
 * mpe_b200.h — C ABI of the B200-native hot path of rpg_monocular_pose_estimator.
 *
 * The reference has no FFI: its boundary is the C++ class API of monocular_pose_estimator_lib
 * (LEDDetector / PoseEstimator / P3P), consumed by MPENode (monocular_pose_estimator/src/
 * monocular_pose_estimator.cpp:84,110-120,159,163-164,222-233).  This header is the thin C boundary that a
 * drop-in replacement of that library binds to (see INTEGRATION.md for the C++ shim that re-creates the
 * class API on top of these entry points).  Each entry point cites the reference interface it replaces;
 * L/ = monocular_pose_estimator_lib/.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types; all pointers are HOST pointers unless the
 *     name says _device.
 *   - return value: 0 = OK, negative = error (MPE_E_*); mpe_last_error(ctx) gives the message.  "No pose
 *     found" / "too few LEDs" are NOT errors (the reference returns false / 0 there): they are reported
 *     through output flags.
 *   - 4x4 poses are ROW-MAJOR double[16] (Eigen is column-major: the C++ shim transposes).
 *   - correspondences are (LED index, detection index) pairs, 1-based, 0 = none — the reference's
 *     VectorXuPairs (L/include/.../datatypes.h:47).
 *   - a context is single-caller (like the reference's PoseEstimator it is not thread-safe); one context
 *     per GPU.
 */
#ifndef MPE_B200_H_
#define MPE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPE_MAX_LEDS 16      /* markers on the object (n_obj)                               */
#define MPE_MAX_DET 16       /* detections that may enter the pose path (n_det)             */
#define MPE_MAX_BLOBS 64     /* detections reported by the LED detector for one frame       */
#define MPE_MAX_DIST 12      /* distortion coefficients accepted (4, 5, 8 or 12; no tilt)   */

enum {
  MPE_OK = 0,
  MPE_E_INVALID = -1,        /* bad argument                                                */
  MPE_E_CUDA = -2,           /* CUDA runtime/driver error (see mpe_last_error)              */
  MPE_E_CAPACITY = -3,       /* batch / image larger than the context was created for       */
  MPE_E_UNSUPPORTED = -4,    /* e.g. gaussian_sigma whose kernel radius exceeds the build's */
  MPE_E_NOT_CONFIGURED = -5  /* camera / markers / params missing                           */
};

/* per-frame flags in mpe_result.flags / mpe_find_leds flags_out */
enum {
  MPE_F_BLOB_OVERFLOW = 1,   /* more than MPE_MAX_BLOBS detections survived the filter      */
  MPE_F_TRACE_ABORT = 2,     /* a contour exceeded the border-follower step limit           */
  MPE_F_TOO_MANY_DET = 4,    /* n_det > MPE_MAX_DET: pose path skipped                      */
  MPE_F_INITIALISED = 8,     /* brute-force initialise() was run for this frame             */
  MPE_F_FULL_IMAGE_RETRY = 16 /* tracking: ROI search failed, whole image was searched      */
};

/* The 11 dynamic-reconfigure tunables (monocular_pose_estimator/cfg/MonocularPoseEstimator.cfg:12-22),
 * i.e. PoseEstimator's public fields (L/include/.../pose_estimator.h:82-91) and its four
 * tolerance/threshold setters (pose_estimator.h, pose_estimator.cpp:187-225). */
typedef struct mpe_params {
  int32_t threshold_value;
  int32_t roi_border_thickness;
  double gaussian_sigma;
  double min_blob_area;
  double max_blob_area;
  double max_width_height_distortion;
  double max_circular_distortion;
  double back_projection_pixel_tolerance;
  double nearest_neighbour_pixel_tolerance;
  double certainty_threshold;
  double valid_correspondence_threshold;
} mpe_params;

typedef struct mpe_rect { int32_t x, y, width, height; } mpe_rect;   /* cv::Rect */

/* One frame's outcome of estimateBodyPose (pose_estimator.cpp:62-147) in the batch entry points. */
typedef struct mpe_result {
  int32_t updated;                         /* pose_updated_ (bool return of estimateBodyPose)        */
  int32_t n_det;                           /* detections found by findLeds                           */
  int32_t n_corr;                          /* rows of correspondences_                               */
  int32_t gn_iters;                        /* Gauss-Newton iterations run by optimisePose            */
  int32_t flags;                           /* MPE_F_*                                                */
  int32_t init_ok;                         /* return of initialise()/checkCorrespondences (0/1)      */
  mpe_rect roi;                            /* region_of_interest_ used for the (last) findLeds       */
  double pose[16];                         /* getPredictedPose(), row-major                          */
  double cov[36];                          /* getPoseCovariance(), row-major 6x6                     */
  uint32_t corr[2 * MPE_MAX_LEDS];         /* correspondences_ rows (LED, detection), 1-based        */
  double det[2 * MPE_MAX_DET];             /* undistorted detections (List2DPoints pixel_positions)  */
  float centers[2 * MPE_MAX_DET];          /* distorted_detection_centers_                           */
} mpe_result;

typedef struct mpe_ctx mpe_ctx;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* Creates a context on CUDA device `device` able to hold `max_batch` frames of at most max_width x
 * max_height pixels.  Replaces PoseEstimator::PoseEstimator() (pose_estimator.cpp:34-42): same defaults
 * (back-projection tol 3, NN tol 5, certainty 0.75, valid-correspondence 0.7). */
int mpe_create(mpe_ctx** out, int device, int max_batch, int max_width, int max_height);
void mpe_destroy(mpe_ctx* ctx);
const char* mpe_last_error(const mpe_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*); NULL restores
 * the context's own stream.  Lets a host framework time the kernels with its own events. */
int mpe_set_stream(mpe_ctx* ctx, void* cuda_stream);

/* ---- configuration ------------------------------------------------------------------------------- */
/* camera_matrix_K_ (row-major 3x3) and camera_distortion_coeffs_ (pose_estimator.h:82-83; written by
 * MPENode::cameraInfoCallback, monocular_pose_estimator.cpp:103-126). nD in {0,4,5,8,12}. */
int mpe_set_camera(mpe_ctx* ctx, const double K[9], const double* D, int nD);
/* PoseEstimator::setMarkerPositions (pose_estimator.cpp:50-55): n x 3 object-frame xyz; also sets
 * histogram_threshold_ = numCombinations(n,3) with the reference's unsigned-factorial arithmetic. */
int mpe_set_markers(mpe_ctx* ctx, const double* xyz, int n);
int mpe_set_params(mpe_ctx* ctx, const mpe_params* p);
/* PoseEstimator::setHistogramThreshold / getHistogramThreshold (pose_estimator.cpp:222-230) */
int mpe_set_histogram_threshold(mpe_ctx* ctx, uint32_t threshold);
uint32_t mpe_get_histogram_threshold(const mpe_ctx* ctx);

/* ---- stage calls (one frame; each mirrors one reference method so it can be parity-tested alone) -- */

/* LEDDetector::findLeds (L/include/.../led_detector.h:84-88, led_detector.cpp:35-112) on the ROI of one
 * 8-bit image (pitch in bytes).  px_out: n x 2 undistorted pixel positions (double, float32-valued);
 * centers_out: n x 2 distorted centres (float).  Capacity of both: MPE_MAX_BLOBS points.  When nothing is
 * found *n_out = 0 (the reference leaves pixel_positions untouched, led_detector.cpp:91 — the shim keeps
 * that quirk).  Uses the context's threshold/sigma/blob parameters and camera. */
int mpe_find_leds(mpe_ctx* ctx, const uint8_t* image, int pitch, int width, int height, mpe_rect roi,
                  double* px_out, float* centers_out, int* n_out, int* flags_out);

/* PoseEstimator::setImagePoints + initialise (pose_estimator.cpp:166-170, 544-721): brute-force P3P
 * correspondence search over det (n_det x 2).  hist_out (optional): n_det x n_obj votes, row = detection.
 * corr_out: k x 2 decoded correspondences (capacity MPE_MAX_LEDS rows), pose_out: the Kabsch pose of
 * checkCorrespondences when *ok = 1. */
int mpe_initialise(mpe_ctx* ctx, const double* det, int n_det, uint32_t* hist_out, uint32_t* corr_out,
                   int* k_out, double pose_out[16], int* ok);

/* PoseEstimator::checkCorrespondences (pose_estimator.cpp:394-542) for given correspondences. */
int mpe_check_correspondences(mpe_ctx* ctx, const double* det, int n_det, const uint32_t* corr, int k,
                              double pose_out[16], int* ok);

/* PoseEstimator::optimisePose (pose_estimator.cpp:733-792): Gauss-Newton from pose_io, returns the
 * refined pose, A^-1 of the last iteration and the iteration count. */
int mpe_optimise_pose(mpe_ctx* ctx, const double* det, int n_det, const uint32_t* corr, int k,
                      double pose_io[16], double cov_out[36], int* iters_out);

/* P3P::computePoses (L/include/.../p3p.h:110-111, p3p.cpp:65-236) for n independent problems.
 * feature_vectors / world_points: n x 9, the three COLUMNS stored one after another (f[9*i + 3*k + r]).
 * solutions: n x 4 x 12 (3x4 row-major [R|C]); status: n ints (0 or -1 colinear). */
int mpe_p3p_compute_poses(mpe_ctx* ctx, const double* feature_vectors, const double* world_points, int n,
                          double* solutions, int* status);

/* ---- batch entry points (the throughput path) ---------------------------------------------------- */

/* Cold mode: every frame is treated as uninitialised (pose_estimator.cpp:68-96): whole-image findLeds,
 * initialise, optimiseAndUpdatePose.  frames: n_frames images of width x height, `pitch` bytes per row,
 * `frame_stride` bytes between frames, in HOST memory (pinned for full PCIe rate).  The call copies the
 * frames to the GPU in chunks overlapped with compute, and returns when results[0..n_frames) are valid. */
int mpe_estimate_batch(mpe_ctx* ctx, const uint8_t* frames, int pitch, long long frame_stride, int width,
                       int height, int n_frames, mpe_result* results);

/* Same, frames already resident in device memory (pitch % 16 == 0, base 16-byte aligned); results is a
 * HOST buffer.  n_frames <= max_batch.  Asynchronous variant: enqueue only, results valid after
 * mpe_synchronize(). */
int mpe_estimate_batch_device(mpe_ctx* ctx, const uint8_t* frames_device, int pitch, long long frame_stride,
                              int width, int height, int n_frames, mpe_result* results);
int mpe_estimate_batch_device_async(mpe_ctx* ctx, const uint8_t* frames_device, int pitch,
                                    long long frame_stride, int width, int height, int n_frames);
int mpe_fetch_results(mpe_ctx* ctx, int n_frames, mpe_result* results);   /* D2H + sync of the last async batch */
int mpe_synchronize(mpe_ctx* ctx);
/* Copies the n poses (row-major 4x4, 16 doubles each) of the last batch into a DEVICE buffer on the context's
 * stream — the record a multi-GPU caller all-gathers over NCCL. */
int mpe_copy_poses_device(mpe_ctx* ctx, int n_frames, double* poses_device);
/* The same for the complete mpe_result records of the last batch / tracking step (n x sizeof(mpe_result) bytes). */
int mpe_copy_results_device(mpe_ctx* ctx, int n_frames, mpe_result* results_device);
/* Host-side helpers of tracking mode (no context, no GPU).  The device loop (mpe_streams_step*) runs these functions per
 * stream on the GPU; callers that drive the stages themselves — PoseEstimator::predictPose / predictMarkerPositionsInImage /
 * predictWithROI of the C++ shim and of the Python mirror — get the same arithmetic here, in the reference's operation order:
 * predictPose (pose_estimator.cpp:232-244: general 4x4 inverse, logarithmMap :996-1064, exponentialMap :962-994),
 * project2d ((K|0)*T first, then the point; :251-268), LEDDetector::determineROI with distortPoints (led_detector.cpp:114-224).
 * Poses are row-major 4x4, pixels n x 2, markers n x 3. */
int mpe_host_predict_pose(const double previous_pose[16], const double current_pose[16], double previous_time, double current_time,
                          double time_to_predict, double predicted_pose_out[16]);
int mpe_host_project_markers(const double K[9], const double pose[16], const double* markers_xyz, int n, double* pixels_out);
int mpe_host_determine_roi(const double* pixels, int n, int width, int height, int border, const double K[9], const double* D, int nD,
                           mpe_rect* roi_out);
int mpe_host_exponential_map(const double twist[6], double pose_out[16]);
int mpe_host_logarithm_map(const double pose[16], double twist_out[6]);
/* Debug aid, host only (no context, no GPU): the 8-bit fixed-point Gaussian taps (sum 256) that mpe_set_params derives from
 * gaussian_sigma — OpenCV's bit-exact kernel for CV_8U with ksize = (0,0) (led_detector.cpp:48-51).  MPE_E_INVALID when the
 * radius would exceed 18 (sigma > ~6.08), MPE_E_CAPACITY when taps_out is too small. */
int mpe_debug_gaussian_taps(double sigma, int* radius_out, uint32_t* taps_out, int taps_capacity);
/* Measurement aid: FP64 throughput of this device (TFLOP/s, FMA = 2) from a kernel of independent DFMA chains — the
 * denominator of the K2 / K3 FP64 roofline in bench.py (MEASURED_PEAKS.json holds no FP64 figure). */
int mpe_probe_fp64_peak(mpe_ctx* ctx, double* tflops_out);

/* Tracking mode: S independent streams, each a PoseEstimator with device-resident state
 * (current/previous/predicted pose, times, it_since_initialized_).  One call advances every stream by
 * one frame exactly as estimateBodyPose does (predictWithROI, ROI findLeds, findCorrespondences,
 * checkCorrespondences, fall back to initialise, whole-image retry; pose_estimator.cpp:97-144).
 * frames_device: one frame per stream, any DEVICE-ACCESSIBLE memory — device memory, or page-locked host memory passed by
 * its device alias (zero-copy: the kernels then read only the ROI tiles over PCIe); times: HOST array of n_streams time
 * stamps; results: HOST array or NULL (then fetch them later with mpe_fetch_results). */
int mpe_streams_reset(mpe_ctx* ctx, int n_streams);
/* Optional indirection for the next mpe_streams_step_device calls: stream s reads image frame_index_device[s] of the buffer
 * (a DEVICE array of n_streams ints; NULL = stream s reads image s).  n_frames_in_buffer = images addressable from the
 * frames pointer.  Lets many streams replay a smaller set of recorded sequences. */
int mpe_streams_set_frame_map(mpe_ctx* ctx, const int* frame_index_device, int n_frames_in_buffer);
int mpe_streams_step_device(mpe_ctx* ctx, const uint8_t* frames_device, int pitch, long long frame_stride,
                            int width, int height, int n_streams, const double* times, mpe_result* results);

/* Host-image variant of the tracking step — the per-image call of a camera driver (MPENode::imageCallback ->
 * PoseEstimator::estimateBodyPose, monocular_pose_estimator.cpp:133-159): frames are n_streams images in HOST memory
 * (image s belongs to stream s; `pitch` bytes per row, `frame_stride` bytes between images), copied to the GPU inside
 * the call; results (HOST, required) are valid on return.  With n_streams = 1 this is a drop-in estimateBodyPose whose
 * PoseEstimator state lives on the GPU.  From its second use with unchanged geometry/configuration the step is replayed
 * as ONE CUDA graph launch (about 30 kernel/memset nodes); mpe_set_graph_replay(ctx, 0) forces plain launches. */
int mpe_streams_step(mpe_ctx* ctx, const uint8_t* frames, int pitch, long long frame_stride, int width, int height,
                     int n_streams, const double* times, mpe_result* results);
int mpe_set_graph_replay(mpe_ctx* ctx, int on);

/* Image ingest of mpe_streams_step (the step BEFORE the hot path: cv_bridge::toCvCopy + the image hand-over of
 * MPENode::imageCallback, monocular_pose_estimator.cpp:143-159).
 *   MPE_INGEST_COPY       every image is copied to the GPU in full (one bulk H2D copy), then searched;
 *   MPE_INGEST_ZERO_COPY  the images must be page-locked (cudaHostAlloc / cudaHostRegister; 16-byte aligned, pitch and
 *                         frame stride multiples of 16): the findLeds kernels read their ROI tiles in place over PCIe
 *                         (TMA from host memory), so in tracking mode only ~1/7 of the image bytes leave the host;
 *   MPE_INGEST_AUTO       (default) zero copy when the images qualify and every stream is tracking, bulk copy otherwise
 *                         (whole-image searches are cheaper through one copy).
 * Graph replay needs a stable image address (e.g. one slot of a pinned ring). */
enum { MPE_INGEST_COPY = 0, MPE_INGEST_ZERO_COPY = 1, MPE_INGEST_AUTO = 2 };
int mpe_set_ingest_mode(mpe_ctx* ctx, int mode);
int mpe_get_ingest_stats(const mpe_ctx* ctx, long long* copy_steps, long long* zero_copy_steps, long long* h2d_bytes_copied);

/* ---- instrumentation ----------------------------------------------------------------------------- */
/* When enabled, the batch entry points bracket each kernel with CUDA events on the launching stream.
 * mpe_get_kernel_times returns, for the last synchronised batch, milliseconds per stage:
 * [0] scan (K1a, the streaming pass), [1] extract_blobs (K1b), [2] p3p_sweep (K2), [3] check + refine (K3a+K3b),
 * [4] blur_tiles (K1c, exact fixed-point blur of the hot tiles). */
int mpe_enable_kernel_timing(mpe_ctx* ctx, int on);
int mpe_get_kernel_times(mpe_ctx* ctx, float ms_out[5]);
/* The brute-force sweep of initialise() (pose_estimator.cpp:565-702) spends its time on hypotheses that never vote.  mode 2
 * (default): a cheap conservative pre-test (csrc/p3p_tier1.cuh) runs in front of the exact P3P solve and only the problems it
 * cannot rule out take the reference's arithmetic; mode 1: every problem is solved exactly and a conservative projection test
 * precedes the exact scoring (the round-1 kernel); mode 0: every finite hypothesis goes through the exact scoring — the
 * all-exact arm for A/B verification.  The histogram is the same in all modes (tests compare them). */
int mpe_set_k2_filter(mpe_ctx* ctx, int mode);
/* number of kernel launches issued by this context so far */
long long mpe_kernel_launch_count(const mpe_ctx* ctx);

/* ---- after the path ------------------------------------------------------------------------------------ */
/* Host-only helper for callers without Eigen: the geometry_msgs/PoseWithCovariance fields MPENode::imageCallback fills from
 * getPredictedPose() / getPoseCovariance() (monocular_pose_estimator.cpp:160-190): position = translation, orientation =
 * Eigen::Quaterniond(R) as (x, y, z, w), covariance elems[j + 6*i] = cov(i, j).  Any output pointer may be NULL. */
void mpe_pose_to_message(const double pose[16], const double cov[36], double position[3], double orientation_xyzw[4],
                         double covariance[36]);

#ifdef __cplusplus
}
#endif
#endif /* MPE_B200_H_ */
