// Internal declarations shared by the kernels and the C-ABI implementation (not installed).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <utility>
#include "../../include/mpe_b200.h"

namespace mpe {

constexpr int kTileRows = 32;        // output rows per K1 tile (= bits of one row-flag word)
constexpr int kMaxRadius = 18;       // largest Gaussian radius (sigma 6 -> ksize 37, the reference's dynamic_reconfigure maximum)
constexpr int kMaxUnrolledRadius = 4; // radii with a fully unrolled blur instantiation; larger ones run the generic loop
constexpr int kMaxTaps = 2 * kMaxRadius + 1;
constexpr int kMaxTileWidthPx = 960; // column-tile width limit (TMA box <= 256 u32 elements incl. halo)
constexpr int kK1Threads = 256;
constexpr int kCandCap = 256;        // candidate contour starts buffered per frame-warp in K1b
constexpr int kMaxFlagWords = 512;   // row-flag words (strips x column tiles) K1b caches per frame; larger geometries are rejected
constexpr int kMaxCombos = MPE_MAX_DET * (MPE_MAX_DET - 1) * (MPE_MAX_DET - 2) / 6;   // 3-subsets of the detections (560)
constexpr int kMaxPerms = MPE_MAX_LEDS * (MPE_MAX_LEDS - 1) * (MPE_MAX_LEDS - 2);     // ordered LED triples (3360)
constexpr int kComboFields = 23;     // K2 detection-triple record (doubles): frame T, f_1, f_2, b, code, packed indices, K T^T
constexpr int kTripleXN = 17;        // first field of the unused-LED coordinates in the LED-triple table
constexpr int kTripleFields = kTripleXN + 3 * (MPE_MAX_LEDS - 3);   // K2 LED-triple table fields (doubles)
constexpr int kK2Queue = 192;        // hypotheses a K2 CTA can park for exact scoring before it scores in place
constexpr int kK2Survivors = 768;    // problems a K2 CTA can park between tier 1 and the exact solve

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: what has been requested is remembered per device
// (one static cache per kernel instantiation), so that a context on a second device of the same process gets its own call.
constexpr int kMaxDevices = 64;
struct SmemAttrCache { std::atomic<size_t> bytes[kMaxDevices]; };
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, SmemAttrCache& cache, size_t default_limit = 48 * 1024) {
  if (bytes <= default_limit) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && bytes <= cache.bytes[dev].load(std::memory_order_acquire)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  if (tracked) {
    size_t cur = cache.bytes[dev].load(std::memory_order_relaxed);
    while (cur < bytes && !cache.bytes[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
  }
  return cudaSuccess;
}

// ---- programmatic dependent launch for the latency path -------------------------------------------------------------------
// A tracking step of a handful of cameras is ~10 tiny kernels, each a single dependent chain; between two of them the GPU
// otherwise drains, then fetches and launches the next grid.  Launched with the programmatic-serialisation attribute the next
// grid may be scheduled while its predecessor winds down; it blocks in pdl_enter() until the predecessor has COMPLETED and its
// writes are visible, so the semantics are those of plain stream order.  Measured, one camera, graph replay: 167.9 -> 165.2 us
// per image; plain launches 184 -> 170 us.  Releasing the dependents at kernel entry instead (MPE_PDL_EARLY=1, all grids of the
// step resident at once) was SLOWER than no attribute at all (171.3 us), so the implicit trigger at grid exit is used.  Rules:
//  * every kernel launched through launch_k() executes pdl_enter() before anything else, in every thread (a grid that finished
//    without waiting would release its successor early);
//  * the attribute is used for small launches only (tl_pdl): a big grid parked on the SMs would take slots from the running one.
extern thread_local bool tl_pdl;
struct PdlScope {
  bool saved;
  explicit PdlScope(bool on) : saved(tl_pdl) { tl_pdl = on; }
  ~PdlScope() { tl_pdl = saved; }
};
#ifdef __CUDACC__
#ifndef MPE_PDL_EARLY
#define MPE_PDL_EARLY 0
#endif
__device__ __forceinline__ void pdl_enter() {
#if MPE_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = tl_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// Camera model as the kernels consume it.
struct DevCamera {
  double K[9];                 // row-major camera_matrix_K_
  double D[MPE_MAX_DIST];      // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (zero padded)
  int nD;
};

struct DevBlobParams {
  double min_blob_area, max_blob_area, max_width_height_distortion, max_circular_distortion;
};

struct DevPoseParams {
  double back_projection_pixel_tolerance;
  double nearest_neighbour_pixel_tolerance;
  double certainty_threshold;
  double valid_correspondence_threshold;
  double back_proj_sq_max;     // largest x with sqrt(x) < back_projection_pixel_tolerance (exact stand-in for the sqrt)
  uint32_t histogram_threshold;
  int n_obj;
  double markers[3 * MPE_MAX_LEDS];
};

constexpr int kFlagNeedsFullStep = 1 << 30;   // internal record flag of the short tracking step, never returned to callers
struct Roi { int x, y, w, h; };   // same layout as mpe_rect / cv::Rect

// Geometry of one K1 launch (uniform over the batch unless `rois` is given).
struct K1Geom {
  int n_frames;
  int img_w, img_h;          // full frame size in pixels
  Roi roi;                   // used when rois == nullptr
  const Roi* rois;           // optional per-frame ROI (device)
  int max_roi_w, max_roi_h;  // bounds for tile enumeration
  int n_strips, n_ct;        // tiles per frame = n_strips * n_ct
  int tw_px;                 // column-tile width in pixels (multiple of 32 when n_ct > 1)
  int box_w;                 // TMA box inner extent in u32 elements (multiple of 4)
  int mask_wpr;              // u32 words per mask row
  int mask_rows;             // rows reserved per frame in the mask buffer
  int flags_per_frame;       // row-flag words reserved per frame
  const int* frame_map;      // optional [n_frames]: index of the image (z coordinate / frame_stride multiple) each entry reads
  // per-frame-ROI launches: the tiles that intersect an ROI, listed by build_tile_list_kernel (f << 12 | strip << 5 | column tile),
  // instead of walking the n_frames x n_strips x n_ct grid and testing every tile against its frame's ROI
  const uint32_t* tile_list; // optional
  const uint32_t* tile_count;
};

struct K1aArgs {
  K1Geom g;
  int thr_k;                 // SWAR constant for "any byte > threshold"
  int threshold;
  int radius;                // Gaussian radius R (taps = 2R+1)
  uint32_t taps[kMaxTaps];   // 8.8 fixed-point Gaussian taps, sum = 256
  uint32_t* rowflags;        // [n_frames][flags_per_frame]   zeroed before the launch; K1c writes the hot tiles' words
  uint32_t* mask;            // [n_frames][mask_rows][mask_wpr]
  // K1a -> K1c hand-over
  uint4* hot_tiles;          // [n_tiles] records {tile id, #hot words (0xffffffff = dense), pool offset, -}
  uint16_t* pool;            // hot words (row << 8 | word) of all hot tiles
  uint32_t pool_capacity;
  uint32_t* counters;        // [0] #hot tiles, [1] pool fill; zeroed before the launch
  const uint8_t* frames;     // frame 0 of this launch (device), for K1c's direct pixel reads
  int pitch;
  long long frame_stride;
};

struct K1bArgs {
  K1Geom g;
  const uint32_t* rowflags;
  const uint32_t* mask;
  DevCamera cam;
  DevBlobParams bp;
  // outputs
  int* n_det;                // [n_frames]
  int* flags;                // [n_frames]
  double* det;               // [n_frames][MPE_MAX_BLOBS][2] undistorted
  float* centers;            // [n_frames][MPE_MAX_BLOBS][2] distorted
  const uint8_t* active;     // optional [n_frames]: frames with 0 are left untouched
};

struct K2Args {
  int n_frames;
  const int* n_det;          // [n_frames]
  const double* det;         // [n_frames][det_stride][2]
  int det_stride;            // points reserved per frame in det
  DevCamera cam;
  DevPoseParams pp;
  int split;                 // CTAs per frame
  uint32_t* hist;            // [n_frames][MPE_MAX_DET*MPE_MAX_LEDS], zeroed before launch
  double* combos;            // [n_frames][kMaxCombos][kComboFields] bearings-only part of every detection triple (prologue kernel)
  const double* triples;     // [kTripleFields][n_perm] LED-triple table (mpe_set_markers)
  double filter_r;           // back-projection tolerance + margin: radius of the conservative reject filters
  int use_filter;            // 0: every hypothesis takes the reference's exact scoring; 1: exact solve + reject filter (round-1 kernel);
                             // 2: tier 1 in front of the exact solve (p3p_tier1.cuh)
  int group;                 // frames a CTA flattens into one problem sequence (tier-1 kernel, split == 1)
  uint32_t* corr;            // [n_frames][2*MPE_MAX_LEDS]
  int* n_corr;               // [n_frames]
  int* frame_flags;          // [n_frames] in/out
  const uint8_t* active;     // optional [n_frames]: run the sweep only where nonzero
  // with `active`: the frames that take part, compacted by compact_active_kernel, so that the sweep's CTAs walk a short list
  // instead of 8192 CTAs finding out one by one that their frame is idle (tracking steps re-initialise ~0.5 % of the streams)
  int* frame_list;           // [n_frames]
  uint32_t* frame_count;     // [1]
};

struct K3Args {
  int n_frames;
  const int* n_det;
  const double* det;
  int det_stride;
  DevCamera cam;
  DevPoseParams pp;
  const uint32_t* corr;      // [n_frames][2*MPE_MAX_LEDS]
  const int* n_corr;         // [n_frames]
  int mode;                  // 0: check + GN (cold path), 1: check only, 2: GN only (pose_io is the start)
  double* pose_io;           // [n_frames][16] row-major
  double* cov;               // [n_frames][36]
  int* ok;                   // [n_frames] checkCorrespondences result
  int* iters;                // [n_frames]
  int* updated;              // [n_frames]
  const uint8_t* active;     // optional
  double* check_sums;        // [n_frames][MPE_MAX_LEDS*3]  sum of H^-1 X_j over the valid subsets (K3a -> K3b)
  int* check_cnt;            // [n_frames][2]  number of valid subsets, number of subsets
  // tracking loop (mode 1 only): what follows from the outcome of checkCorrespondences, written per frame by refine_kernel
  uint8_t* set_gn_if_ok;     // optional [n_frames]: set to 1 when the check succeeded (-> optimisePose)
  uint8_t* set_init_if_fail; // optional [n_frames]: set to 1 when it failed (-> brute-force initialise(), pose_estimator.cpp:842)
};

// Per-stream state of the tracking loop = the PoseEstimator members that survive a frame (pose_estimator.h:56-79)
struct StreamState {
  double current_pose[16], previous_pose[16], predicted_pose[16];   // row-major
  double current_time, previous_time, predicted_time;
  int it_since_initialized;
  int pad;
};

struct TrackArgs {
  int n;
  int img_w, img_h, roi_border;
  StreamState* state;
  const double* times;       // [n] time_to_predict of this step
  DevCamera cam;
  DevPoseParams pp;
  Roi* rois;                 // [n] ROI handed to K1 (empty = skip)
  Roi* result_rois;          // [n] region_of_interest_ reported for the frame
  double* pred_px;           // [n][MPE_MAX_LEDS][2] predicted_pixel_positions_
  uint8_t *mode, *done, *a_retry, *a_check, *a_init, *a_gn;   // [n] each
  int* track_flags;          // [n]
  int fast;                  // 1: the short step (no whole-image retry, no re-initialisation stages): a stream that would need them is
                             //    left untouched and flagged kFlagNeedsFullStep, the host then runs the complete step
  // buffers shared with the cold path
  const int* n_det; const int* flags; const double* det; const float* centers;
  uint32_t* corr; int* n_corr; double* pose_io; const double* cov; int* ok; int* iters; int* updated;
};

// ---- launchers (defined next to the kernels) ----
cudaError_t launch_track_begin(const TrackArgs& a, cudaStream_t st);
cudaError_t launch_track_after_detect(const TrackArgs& a, int pass, cudaStream_t st);
cudaError_t launch_track_finish(const TrackArgs& a, mpe_result* out, mpe_result* out_host, cudaStream_t st);
cudaError_t launch_track_reset(StreamState* s, int n, cudaStream_t st);
cudaError_t launch_find_leds(const K1aArgs& a, const CUtensorMap& tmap, int radius, int n_sms, cudaStream_t st);
cudaError_t launch_build_tile_list(const K1Geom& g, uint32_t* list, uint32_t* count, cudaStream_t st);
cudaError_t launch_blur_tiles(const K1aArgs& a, int radius, int n_sms, cudaStream_t st);
cudaError_t launch_extract_blobs(const K1bArgs& a, cudaStream_t st);
cudaError_t launch_p3p_sweep(const K2Args& a, int n_sms, cudaStream_t st);
cudaError_t launch_marker_triples(const DevPoseParams& pp, double* table, cudaStream_t st);
cudaError_t launch_validate_refine(const K3Args& a, cudaStream_t st);
cudaError_t launch_p3p_batch(const double* f, const double* P, int n, double* sol, int* status, cudaStream_t st);
size_t find_leds_smem_bytes(const K1Geom& g, int radius, int stages);

}  // namespace mpe
