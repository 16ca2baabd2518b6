// C-ABI implementation (include/mpe_b200.h): context, configuration, stage calls and the batch drivers.
// Host code only orchestrates: every arithmetic step of the hot path runs in the CUDA kernels of
// k1_find_leds.cu / k2_p3p_sweep.cu / k3_validate_refine.cu.  There is no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "mpe_internal.cuh"
#include "tracking_math.cuh"

using namespace mpe;

namespace mpe { thread_local bool tl_pdl = false; }

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kTrackTileWidthPx = 256;   // column-tile width of the tracking-mode K1 launches (ROIs are ~100-200 px wide)
constexpr int kLanes = 4;                // sub-batches of one cold batch that may be in flight at once (two streams, alternating)
constexpr int kPipelineMinFrames = 1024; // below this a batch is not split
constexpr size_t kPoolPerFrame = 1024;   // hot-word pool entries reserved per frame of capacity (overflow degrades to the dense path)

struct DevBuffers {
  uint8_t* frames = nullptr;       // [max_batch][max_h][pitch]
  uint32_t* rowflags = nullptr;    // [max_batch][flags_per_frame]
  uint32_t* mask = nullptr;        // [max_batch][max_h][mask_wpr]
  int* n_det = nullptr;            // [max_batch]
  int* flags = nullptr;            // [max_batch]
  double* det = nullptr;           // [max_batch][MPE_MAX_BLOBS][2]
  float* centers = nullptr;        // [max_batch][MPE_MAX_BLOBS][2]
  uint32_t* hist = nullptr;        // [max_batch][MPE_MAX_DET*MPE_MAX_LEDS]
  double* combos = nullptr;        // [max_batch][kMaxCombos][kComboFields]  K2 detection-triple table
  double* triples = nullptr;       // [kTripleFields][kMaxPerms]             K2 LED-triple table (rebuilt by mpe_set_markers)
  uint32_t* corr = nullptr;        // [max_batch][2*MPE_MAX_LEDS]
  int* n_corr = nullptr;           // [max_batch]
  double* pose = nullptr;          // [max_batch][16]
  double* cov = nullptr;           // [max_batch][36]
  int* ok = nullptr;               // [max_batch]
  int* iters = nullptr;            // [max_batch]
  int* updated = nullptr;          // [max_batch]
  Roi* rois = nullptr;             // [max_batch]
  mpe_result* results = nullptr;   // [max_batch]
  double* check_sums = nullptr;    // [max_batch][MPE_MAX_LEDS*3]
  uint4* hot_tiles = nullptr;      // [max_batch * flags_per_frame]
  uint16_t* pool = nullptr;        // [max_batch * kPoolPerFrame]
  uint32_t* counters = nullptr;    // [kLanes][2]  (one pair per concurrently running sub-batch)
  uint32_t* tile_list = nullptr;   // [max_batch * flags_per_frame]  K1 work list of per-stream-ROI launches
  uint32_t* tile_count = nullptr;  // [1]
  int* k2_list = nullptr;          // [max_batch]  frames taking part in a masked K2 sweep
  uint32_t* k2_count = nullptr;    // [1]
  // tracking loop
  StreamState* streams = nullptr;  // [max_batch]
  Roi* result_rois = nullptr;      // [max_batch]
  double* pred_px = nullptr;       // [max_batch][MPE_MAX_LEDS][2]
  uint8_t* masks = nullptr;        // 6 x [max_batch]: mode, done, a_retry, a_check, a_init, a_gn
  int* track_flags = nullptr;      // [max_batch]
  double* times = nullptr;         // [max_batch]
  int* check_cnt = nullptr;        // [max_batch][2]
};

}  // namespace

struct mpe_ctx {
  int device = 0;
  int n_sms = 148;
  int max_batch = 0, max_w = 0, max_h = 0;
  int pitch = 0;                  // device frame pitch (multiple of 16)
  int mask_wpr = 0, flags_per_frame = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pipeline = false;           // split large cold batches over two streams: measured SLOWER (3.06 vs 2.80 ms @8192); MPE_PIPELINE=1 enables
  DevBuffers d;
  mpe_result* h_results = nullptr;   // pinned
  double* h_times = nullptr;         // pinned: time stamps read in place by the tracking step of a few cameras
  mpe_result* h_results_dev = nullptr; const double* h_times_dev = nullptr;   // their device aliases
  // configuration
  bool have_camera = false, have_markers = false, have_params = false;
  DevCamera cam{};
  DevPoseParams pp{};
  mpe_params params{};
  // instrumentation
  bool timing = false;
  bool mapped_io = true;             // MPE_MAPPED_IO=0 switches the in-place time stamps / records of small steps off
  cudaEvent_t ev[10]{};
  bool timing_pending = false;
  long long launches = 0;
  std::string err;
  PFN_encodeTiled encode = nullptr;
  const int* frame_map = nullptr;   // optional stream -> image index table for mpe_streams_step_device
  int frame_map_total = 0;
  std::vector<cudaEvent_t> chunk_events;
  // CUDA-graph replay of the tracking step (mpe_streams_step*): the ~30 small launches of one step become one graph launch
  int k2_filter = 2;                      // K2 reject filters: 0 none (every hypothesis scored exactly; MPE_K2_NO_FILTER=1), 1 after the exact solve, 2 tier 1 in front of it
  bool use_graphs = true;
  unsigned long long cfg_version = 0;     // bumped by every configuration call; a stale graph is rebuilt
  struct StepGraphKey {
    const uint8_t* frames; int pitch; long long stride; int w, h, n; const int* fmap; int fmap_total; int fetch;
    unsigned long long cfg; cudaStream_t st;
    bool operator==(const StepGraphKey& o) const {
      return frames == o.frames && pitch == o.pitch && stride == o.stride && w == o.w && h == o.h && n == o.n && fmap == o.fmap &&
             fmap_total == o.fmap_total && fetch == o.fetch && cfg == o.cfg && st == o.st;
    }
  };
  // two cached graphs: [0] the complete step, [1] the short step (small n, every stream tracking; see run_streams_step)
  StepGraphKey graph_key[2]{}, pending_key[2]{};   // a key is run eagerly once (function attributes get configured), captured on its second use
  bool have_pending[2] = {false, false};
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  bool short_steps = true;                // MPE_SHORT_STEPS=0 disables the short tracking step
  long long short_step_count = 0, short_step_fallbacks = 0;
  // image ingest of the host-image tracking step (mpe_streams_step)
  int ingest_mode = MPE_INGEST_AUTO;
  std::vector<uint8_t> stream_tracking;   // host mirror: stream s has produced a pose since the last reset (it_since_initialized_ >= 1)
  int n_tracking = 0;
  long long h2d_bytes_copied = 0, zero_copy_steps = 0, copy_steps = 0;
  long long graph_launches[2] = {0, 0};   // kernel launches contained in one replay
  long long graph_replays = 0;
};

namespace {

int fail(mpe_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CUDA_TRY(ctx, expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(ctx, MPE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// OpenCV getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED for CV_8U (ksize from sigma:
// cvRound(sigma*6+1)|1), 8 fractional bits, centre tap = 256 - sum(others).  Verified against cv2 4.13
// impulse responses and full-image blurs for sigma in [0.3, 6] (tests/test_cpu_gaussian_taps.py sweeps sigma on the CPU through mpe_debug_gaussian_taps, tests/test_gpu_find_leds.py compares whole frames).
bool gaussian_taps_8u(double sigma, int* radius, uint32_t taps[kMaxTaps]) {
  if (!(sigma > 0)) return false;
  int n = (int)std::nearbyint(sigma * 6 + 1) | 1;
  int n2 = (n - 1) / 2;
  if (n2 > kMaxRadius) return false;
  if (n2 < 1) {            // sigma < 1/12: OpenCV's kernel has one tap, the blur is the identity; same as radius 1 with taps 0 256 0
    taps[0] = 0; taps[1] = 256; taps[2] = 0;
    *radius = 1;
    return true;
  }
  std::vector<double> vals(n2);
  double scale2x = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n2; ++i) {
    double x = (double)(i - n2);
    vals[i] = std::exp(scale2x * x * x);
    sum += vals[i];
  }
  sum = 2 * sum + 1.0;
  double mul = 1.0 / sum;
  double err = 0;
  long long tot = 0;
  for (int i = 0; i < n2; ++i) {
    double adj = vals[i] * mul * 256.0 + err;
    long long v0 = (long long)std::nearbyint(adj);
    err = adj - (double)v0;
    taps[i] = (uint32_t)v0;
    taps[n - 1 - i] = (uint32_t)v0;
    tot += 2 * v0;
  }
  taps[n2] = (uint32_t)(256 - tot);
  *radius = n2;
  return true;
}

// Largest double x such that sqrt(x) < tol in IEEE arithmetic (sqrt is correctly rounded and monotonic), so that the sweep can
// test the squared distance: "sqrt(d2) < tol" (pose_estimator.cpp:671,689)  <=>  "d2 <= x".  -1 when no d2 >= 0 qualifies.
double sqrt_less_than_bound(double tol) {
  if (!(tol > 0)) return -1.0;
  if (std::isinf(tol)) return std::numeric_limits<double>::max();
  double x = tol * tol;
  while (std::sqrt(x) >= tol) x = std::nextafter(x, 0.0);
  while (std::sqrt(std::nextafter(x, INFINITY)) < tol) x = std::nextafter(x, INFINITY);
  return x;
}

// Combinations::numCombinations with the reference's unsigned factorial (combinations.cpp:34-45)
unsigned factorial_u(int N) { return (N == 1 || N == 0) ? 1u : factorial_u(N - 1) * (unsigned)N; }
unsigned num_combinations_ref(unsigned N, unsigned K) { return factorial_u((int)N) / (factorial_u((int)K) * factorial_u((int)(N - K))); }

struct FrameSource {
  const uint8_t* base;      // device pointer of frame 0
  int pitch;                // bytes per row (multiple of 16)
  long long frame_stride;   // bytes between frames (multiple of 16)
  int width, height;
  int n_frames_total;       // frames addressable from base
};

int make_geometry(mpe_ctx* c, int w, int h, Roi roi, int max_roi_w, int max_roi_h, int radius, int forced_tw, K1Geom* g) {
  g->img_w = w; g->img_h = h;
  g->roi = roi; g->rois = nullptr;
  g->max_roi_w = max_roi_w; g->max_roi_h = max_roi_h;
  g->n_strips = (max_roi_h + kTileRows - 1) / kTileRows;
  int tw = max_roi_w;
  if (tw > kMaxTileWidthPx) tw = kMaxTileWidthPx;          // 960 = 30 mask words
  if (forced_tw > 0 && forced_tw < tw) tw = forced_tw;     // tracking: narrow column tiles (multiple of 32) for small ROIs
  g->n_ct = (max_roi_w + tw - 1) / tw;
  if (g->n_ct > 1 && forced_tw <= 0) tw = kMaxTileWidthPx;
  g->tw_px = tw;
  g->frame_map = nullptr;
  g->tile_list = nullptr; g->tile_count = nullptr;
  // widest span of u32 elements a tile can touch: it starts at the 16-pixel boundary at or below x - R (TMA needs a
  // 16-byte aligned innermost coordinate) and must reach pixel x + tw + R - 1
  int span = (tw + 2 * radius - 1 + 15) / 4 + 1;
  g->box_w = (span + 3) & ~3;
  if (g->box_w > 256) return MPE_E_UNSUPPORTED;
  g->mask_wpr = c->mask_wpr;
  g->mask_rows = c->max_h;
  g->flags_per_frame = c->flags_per_frame;
  if (g->n_strips * g->n_ct > g->flags_per_frame) return MPE_E_CAPACITY;
  if (g->n_strips * g->n_ct > kMaxFlagWords) return MPE_E_CAPACITY;      // K1b's per-frame cache of row flags
  if (g->n_ct * ((tw + 31) / 32) > g->mask_wpr && g->n_ct > 1) return MPE_E_CAPACITY;
  return MPE_OK;
}

int encode_tensor_map(mpe_ctx* c, const FrameSource& src, int box_w, int rows, CUtensorMap* out) {
  if (!c->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) return fail(c, MPE_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    c->encode = (PFN_encodeTiled)fn;
  }
  if ((src.pitch & 15) || (src.frame_stride & 15) || ((uintptr_t)src.base & 15))
    return fail(c, MPE_E_INVALID, "device frames must be 16-byte aligned with pitch and frame stride multiples of 16");
  cuuint64_t gdim[3] = {(cuuint64_t)(src.pitch / 4), (cuuint64_t)src.height, (cuuint64_t)src.n_frames_total};
  cuuint64_t gstride[2] = {(cuuint64_t)src.pitch, (cuuint64_t)src.frame_stride};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = c->encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)src.base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, MPE_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return MPE_OK;
}

// MPE_STEP_TRACE=1 (diagnostic, plain launches only): CUDA events between the stages of the tracking step; the previous step's
// stage times are printed to stderr when the next step starts.
struct StepTrace {
  static constexpr int kMax = 32;
  cudaEvent_t ev[kMax]; const char* name[kMax]; int n = 0; bool made = false; int on = -1; bool pending = false;
  bool enabled() { if (on < 0) { const char* e = getenv("MPE_STEP_TRACE"); on = (e && atoi(e)) ? 1 : 0; } return on == 1; }
  void begin(cudaStream_t st) {
    if (!enabled()) return;
    if (!made) { for (auto& e : ev) cudaEventCreate(&e); made = true; }
    if (pending) {
      cudaEventSynchronize(ev[n - 1]);
      fprintf(stderr, "[step trace]");
      for (int i = 1; i < n; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, " %s %.1f", name[i], ms * 1e3f); }
      float tot = 0; cudaEventElapsedTime(&tot, ev[0], ev[n - 1]); fprintf(stderr, " | total %.1f us\n", tot * 1e3f);
    }
    n = 0; pending = true; mark("begin", st);
  }
  void mark(const char* what, cudaStream_t st) { if (!enabled() || n >= kMax) return; name[n] = what; cudaEventRecord(ev[n++], st); }
};
static StepTrace g_trace;

void time_begin(mpe_ctx* c, int k, cudaStream_t st) { if (c->timing) cudaEventRecord(c->ev[2 * k], st); }
void time_end(mpe_ctx* c, int k, cudaStream_t st) {
  if (c->timing) { cudaEventRecord(c->ev[2 * k + 1], st); c->timing_pending = true; }
  static const char* const kStage[5] = {"scan", "extract", "sweep", "check/refine", "blur"};
  if (g_trace.pending) g_trace.mark(kStage[k], st);
}

// K1a + K1b over frames [f0, f0+n) of `src`, outputs into slots [slot0, slot0+n) of the context buffers.
int run_find_leds(mpe_ctx* c, const FrameSource& src, int f0, int n, int slot0, Roi roi, const Roi* rois_dev, cudaStream_t st,
                  int forced_tw = 0, const int* frame_map = nullptr, const uint8_t* active = nullptr, int lane = 0) {
  int radius = 0;
  K1aArgs a{};
  if (!gaussian_taps_8u(c->params.gaussian_sigma, &radius, a.taps))
    return fail(c, MPE_E_UNSUPPORTED, "gaussian_sigma must give a kernel radius in [1, " + std::to_string(kMaxRadius) + "]");
  int max_w = rois_dev ? src.width : roi.w, max_h = rois_dev ? src.height : roi.h;
  int rc = make_geometry(c, src.width, src.height, roi, max_w, max_h, radius, forced_tw, &a.g);
  if (rc != MPE_OK) return fail(c, rc, "ROI / image too large for this context");
  a.g.n_frames = n;
  a.g.rois = rois_dev;
  a.g.frame_map = frame_map;
  // per-frame ROIs: list the tiles the ROIs touch (the packed entry holds 20 bits of frame, 7 of strip, 5 of column tile)
  const bool listed = rois_dev && a.g.n_strips <= 128 && a.g.n_ct <= 32 && n < (1 << 20);
  if (listed) {
    CUDA_TRY(c, cudaMemsetAsync(c->d.tile_count, 0, sizeof(uint32_t), st));
    CUDA_TRY(c, launch_build_tile_list(a.g, c->d.tile_list, c->d.tile_count, st));
    ++c->launches;
    a.g.tile_list = c->d.tile_list;
    a.g.tile_count = c->d.tile_count;
  }
  CUtensorMap tmap;
  FrameSource sub = src;
  sub.base = src.base + (size_t)f0 * src.frame_stride;
  sub.n_frames_total = frame_map ? src.n_frames_total : n;
  rc = encode_tensor_map(c, sub, a.g.box_w, kTileRows + 2 * radius, &tmap);
  if (rc != MPE_OK) return rc;
  int T = c->params.threshold_value;
  if (T < 0) T = 0;                  // THRESH_TOZERO with a negative threshold keeps every pixel; so does "> 0" (0 stays 0)
  if (T > 255) T = 255;              // nothing is > 255
  a.threshold = T;
  a.radius = radius;
  a.thr_k = (T < 128) ? (127 - T) * 0x01010101 : (255 - T) * 0x01010101;
  a.rowflags = c->d.rowflags + (size_t)slot0 * c->flags_per_frame;
  a.mask = c->d.mask + (size_t)slot0 * c->max_h * c->mask_wpr;
  a.hot_tiles = c->d.hot_tiles + (size_t)slot0 * c->flags_per_frame;
  a.pool = c->d.pool + (size_t)slot0 * kPoolPerFrame;
  a.pool_capacity = (uint32_t)((size_t)n * kPoolPerFrame);
  a.counters = c->d.counters + 2 * lane;
  a.frames = sub.base;
  a.pitch = src.pitch;
  a.frame_stride = src.frame_stride;
  CUDA_TRY(c, cudaMemsetAsync(a.rowflags, 0, (size_t)n * c->flags_per_frame * sizeof(uint32_t), st));
  CUDA_TRY(c, cudaMemsetAsync(a.counters, 0, 2 * sizeof(uint32_t), st));
  time_begin(c, 0, st);
  CUDA_TRY(c, launch_find_leds(a, tmap, radius, c->n_sms, st));
  time_end(c, 0, st);
  time_begin(c, 4, st);
  CUDA_TRY(c, launch_blur_tiles(a, radius, c->n_sms, st));
  time_end(c, 4, st);
  c->launches += 2;

  K1bArgs b{};
  b.g = a.g;
  b.rowflags = a.rowflags;
  b.mask = a.mask;
  b.cam = c->cam;
  b.bp.min_blob_area = c->params.min_blob_area;
  b.bp.max_blob_area = c->params.max_blob_area;
  b.bp.max_width_height_distortion = c->params.max_width_height_distortion;
  b.bp.max_circular_distortion = c->params.max_circular_distortion;
  b.n_det = c->d.n_det + slot0;
  b.flags = c->d.flags + slot0;
  b.det = c->d.det + (size_t)slot0 * MPE_MAX_BLOBS * 2;
  b.centers = c->d.centers + (size_t)slot0 * MPE_MAX_BLOBS * 2;
  b.active = active;
  time_begin(c, 1, st);
  CUDA_TRY(c, launch_extract_blobs(b, st));
  time_end(c, 1, st);
  ++c->launches;
  return MPE_OK;
}

int choose_split(const mpe_ctx* c, int n_frames) {
  int n = c->pp.n_obj;
  long long est = (long long)n * (n - 1) * (n - 2) / 6 * n * (n - 1) * (n - 2);
  long long by_work = (est + 255) / 256;
  long long by_fill = (4LL * c->n_sms + n_frames - 1) / n_frames;
  long long s = by_work < by_fill ? by_work : by_fill;
  if (s < 1) s = 1;
  if (s > 128) s = 128;
  return (int)s;
}

int run_sweep(mpe_ctx* c, int slot0, int n, cudaStream_t st, const uint8_t* active) {
  K2Args k{};
  k.n_frames = n;
  k.n_det = c->d.n_det + slot0;
  k.det = c->d.det + (size_t)slot0 * MPE_MAX_BLOBS * 2;
  k.det_stride = MPE_MAX_BLOBS;
  k.cam = c->cam;
  k.pp = c->pp;
  k.split = choose_split(c, n);
  k.hist = c->d.hist + (size_t)slot0 * MPE_MAX_DET * MPE_MAX_LEDS;
  k.combos = c->d.combos + (size_t)slot0 * kMaxCombos * kComboFields;
  k.triples = c->d.triples;
  // conservative reject filter (k2_p3p_sweep.cu: maybe_within): radius = tolerance + margin; trusted only while its error bound
  // 4e-6 * fx pixels is far below the margin
  {
    const double tol = c->pp.back_projection_pixel_tolerance;
    const double fmax_ = std::fmax(std::fabs(c->cam.K[0]), std::fabs(c->cam.K[4]));
    const double margin = 0.25;
    k.use_filter = (c->k2_filter && std::isfinite(tol) && tol > 0 && std::isfinite(fmax_) && 4e-6 * fmax_ * 8 <= margin) ? c->k2_filter : 0;
    k.filter_r = tol + margin;
  }
  k.corr = c->d.corr + (size_t)slot0 * 2 * MPE_MAX_LEDS;
  k.n_corr = c->d.n_corr + slot0;
  k.frame_flags = c->d.flags + slot0;
  k.active = active;
  if (active) {
    k.frame_list = c->d.k2_list;
    k.frame_count = c->d.k2_count;
    CUDA_TRY(c, cudaMemsetAsync(k.frame_count, 0, sizeof(uint32_t), st));
    ++c->launches;                                     // compact_active_kernel
  }
  CUDA_TRY(c, cudaMemsetAsync(k.hist, 0, (size_t)n * MPE_MAX_DET * MPE_MAX_LEDS * sizeof(uint32_t), st));
  time_begin(c, 2, st);
  CUDA_TRY(c, launch_p3p_sweep(k, c->n_sms, st));
  time_end(c, 2, st);
  c->launches += 3;
  return MPE_OK;
}

int run_refine(mpe_ctx* c, int slot0, int n, int mode, cudaStream_t st, const uint8_t* active, uint8_t* set_gn_if_ok = nullptr,
               uint8_t* set_init_if_fail = nullptr) {
  K3Args k{};
  k.n_frames = n;
  k.n_det = c->d.n_det + slot0;
  k.det = c->d.det + (size_t)slot0 * MPE_MAX_BLOBS * 2;
  k.det_stride = MPE_MAX_BLOBS;
  k.cam = c->cam;
  k.pp = c->pp;
  k.corr = c->d.corr + (size_t)slot0 * 2 * MPE_MAX_LEDS;
  k.n_corr = c->d.n_corr + slot0;
  k.mode = mode;
  k.pose_io = c->d.pose + (size_t)slot0 * 16;
  k.cov = c->d.cov + (size_t)slot0 * 36;
  k.ok = c->d.ok + slot0;
  k.iters = c->d.iters + slot0;
  k.updated = c->d.updated + slot0;
  k.active = active;
  k.check_sums = c->d.check_sums + (size_t)slot0 * MPE_MAX_LEDS * 3;
  k.check_cnt = c->d.check_cnt + (size_t)slot0 * 2;
  k.set_gn_if_ok = set_gn_if_ok;
  k.set_init_if_fail = set_init_if_fail;
  time_begin(c, 3, st);
  CUDA_TRY(c, launch_validate_refine(k, st));
  time_end(c, 3, st);
  c->launches += (mode == 2) ? 1 : ((mode == 1 || n > 1024) ? 2 : 3);   // check + Kabsch/GN in one kernel, or check + Kabsch + cooperative GN for small batches
  return MPE_OK;
}

__global__ void pack_results_kernel(int n, Roi roi, const int* n_det, const int* flags, const double* det, const float* centers,
                                    const uint32_t* corr, const int* n_corr, const double* pose, const double* cov, const int* ok,
                                    const int* iters, const int* updated, mpe_result* out) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  mpe_result r;
  r.updated = updated[f];
  r.n_det = n_det[f];
  r.n_corr = n_corr[f];
  r.gn_iters = iters[f];
  r.flags = flags[f];
  r.init_ok = ok[f];
  r.roi.x = roi.x; r.roi.y = roi.y; r.roi.width = roi.w; r.roi.height = roi.h;
  for (int i = 0; i < 16; ++i) r.pose[i] = pose[(size_t)f * 16 + i];
  for (int i = 0; i < 36; ++i) r.cov[i] = cov[(size_t)f * 36 + i];
  for (int i = 0; i < 2 * MPE_MAX_LEDS; ++i) r.corr[i] = (i < 2 * r.n_corr) ? corr[(size_t)f * 2 * MPE_MAX_LEDS + i] : 0u;
  int nd = r.n_det < MPE_MAX_DET ? r.n_det : MPE_MAX_DET;
  for (int i = 0; i < 2 * MPE_MAX_DET; ++i) {
    r.det[i] = (i < 2 * nd) ? det[(size_t)f * MPE_MAX_BLOBS * 2 + i] : 0.0;
    r.centers[i] = (i < 2 * nd) ? centers[(size_t)f * MPE_MAX_BLOBS * 2 + i] : 0.f;
  }
  out[f] = r;
}

// The whole cold path for frames [f0, f0+n) of src -> device results slots [slot0, slot0+n)
int run_cold(mpe_ctx* c, const FrameSource& src, int f0, int n, int slot0, cudaStream_t st, int lane = 0) {
  Roi full{0, 0, src.width, src.height};                       // pose_estimator.cpp:72
  int rc = run_find_leds(c, src, f0, n, slot0, full, nullptr, st, 0, nullptr, nullptr, lane);
  if (rc != MPE_OK) return rc;
  rc = run_sweep(c, slot0, n, st, nullptr);                    // setImagePoints + initialise (histogram + decode)
  if (rc != MPE_OK) return rc;
  CUDA_TRY(c, cudaMemsetAsync(c->d.pose + (size_t)slot0 * 16, 0, (size_t)n * 16 * sizeof(double), st));
  CUDA_TRY(c, cudaMemsetAsync(c->d.cov + (size_t)slot0 * 36, 0, (size_t)n * 36 * sizeof(double), st));
  rc = run_refine(c, slot0, n, 0, st, nullptr);                // checkCorrespondences + optimisePose
  if (rc != MPE_OK) return rc;
  pack_results_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, full, c->d.n_det + slot0, c->d.flags + slot0,
                                                       c->d.det + (size_t)slot0 * MPE_MAX_BLOBS * 2, c->d.centers + (size_t)slot0 * MPE_MAX_BLOBS * 2,
                                                       c->d.corr + (size_t)slot0 * 2 * MPE_MAX_LEDS, c->d.n_corr + slot0, c->d.pose + (size_t)slot0 * 16,
                                                       c->d.cov + (size_t)slot0 * 36, c->d.ok + slot0, c->d.iters + slot0, c->d.updated + slot0,
                                                       c->d.results + slot0);
  CUDA_TRY(c, cudaGetLastError());
  ++c->launches;
  return MPE_OK;
}

// A large cold batch as kLanes sub-batches alternating on two streams: the latency-bound, low-occupancy stages of one
// sub-batch (contour tracing with a warp per frame, Gauss-Newton with a thread per frame, the sparse blur) overlap the wide
// stages of the other (the HBM-bound scan, the FP64-bound P3P sweep).  Sub-batches own disjoint slots of every buffer.
// Per-kernel event timing (mpe_enable_kernel_timing) needs the stages back to back and switches the split off.
// EXPERIMENT, off by default: on B200 the four sets of kernel ramps and tails cost more than the overlap wins (8192 frames:
// 3.06 ms split against 2.80 ms back to back; 1080p: 1.71 against 1.35 ms).
int run_cold_pipelined(mpe_ctx* c, const FrameSource& src, int n, cudaStream_t st) {
  if (!c->pipeline || c->timing || n < kPipelineMinFrames) return run_cold(c, src, 0, n, 0, st);
  const int per = (n + kLanes - 1) / kLanes;
  CUDA_TRY(c, cudaEventRecord(c->ev_fork, st));
  CUDA_TRY(c, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
  for (int p = 0; p < kLanes; ++p) {
    const int f0 = p * per;
    const int np = (n - f0 < per) ? n - f0 : per;
    if (np <= 0) break;
    int rc = run_cold(c, src, f0, np, f0, (p & 1) ? c->aux_stream : st, p);
    if (rc != MPE_OK) return rc;
  }
  CUDA_TRY(c, cudaEventRecord(c->ev_join, c->aux_stream));
  CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_join, 0));
  return MPE_OK;
}

int check_configured(mpe_ctx* c, bool need_markers) {
  if (!c) return MPE_E_INVALID;
  if (!c->have_camera) return fail(c, MPE_E_NOT_CONFIGURED, "camera not set (mpe_set_camera)");
  if (!c->have_params) return fail(c, MPE_E_NOT_CONFIGURED, "parameters not set (mpe_set_params)");
  if (need_markers && !c->have_markers) return fail(c, MPE_E_NOT_CONFIGURED, "markers not set (mpe_set_markers)");
  return MPE_OK;
}

template <typename T>
cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

}  // namespace

// =====================================================================================================
extern "C" {

int mpe_create(mpe_ctx** out, int device, int max_batch, int max_width, int max_height) {
  if (!out || max_batch < 1 || max_width < 8 || max_height < 1) return MPE_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return MPE_E_CUDA;
  mpe_ctx* c = new mpe_ctx();
  c->device = device;
#define CREATE_TRY(expr)                                                           \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      fprintf(stderr, "mpe_create: %s: %s\n", #expr, cudaGetErrorString(_e));      \
      mpe_destroy(c);                                                              \
      return MPE_E_CUDA;                                                           \
    }                                                                              \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  c->n_sms = prop.multiProcessorCount;
  c->max_batch = max_batch; c->max_w = max_width; c->max_h = max_height;
  c->pitch = (max_width + 15) & ~15;
  int n_ct = (max_width + kMaxTileWidthPx - 1) / kMaxTileWidthPx;
  int n_ct_track = (max_width + kTrackTileWidthPx - 1) / kTrackTileWidthPx;      // tracking mode uses narrow column tiles
  c->mask_wpr = (n_ct > 1) ? n_ct * (kMaxTileWidthPx / 32) : (max_width + 31) / 32;
  if (n_ct_track * (kTrackTileWidthPx / 32) > c->mask_wpr) c->mask_wpr = n_ct_track * (kTrackTileWidthPx / 32);
  c->flags_per_frame = ((max_height + kTileRows - 1) / kTileRows) * (n_ct > n_ct_track ? n_ct : n_ct_track);
  CREATE_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  { const char* e = getenv("MPE_PIPELINE"); c->pipeline = (e && e[0] == '1'); }
  c->stream = c->own_stream;
  size_t B = (size_t)max_batch;
  CREATE_TRY(dev_alloc(&c->d.frames, B * c->pitch * max_height));
  CREATE_TRY(dev_alloc(&c->d.rowflags, B * c->flags_per_frame));
  CREATE_TRY(dev_alloc(&c->d.mask, B * max_height * c->mask_wpr));
  CREATE_TRY(dev_alloc(&c->d.n_det, B));
  CREATE_TRY(dev_alloc(&c->d.flags, B));
  CREATE_TRY(dev_alloc(&c->d.det, B * MPE_MAX_BLOBS * 2));
  CREATE_TRY(dev_alloc(&c->d.centers, B * MPE_MAX_BLOBS * 2));
  CREATE_TRY(dev_alloc(&c->d.hist, B * MPE_MAX_DET * MPE_MAX_LEDS));
  CREATE_TRY(dev_alloc(&c->d.combos, B * kMaxCombos * kComboFields));
  CREATE_TRY(dev_alloc(&c->d.triples, (size_t)kTripleFields * kMaxPerms));
  { const char* e = getenv("MPE_K2_NO_FILTER"); if (e && e[0] == '1') c->k2_filter = 0; }
  { const char* e = getenv("MPE_SHORT_STEPS"); if (e && e[0] == '0') c->short_steps = false; }
  { const char* e = getenv("MPE_MAPPED_IO"); if (e && e[0] == '0') c->mapped_io = false; }
  { const char* e = getenv("MPE_K2_FILTER"); if (e && e[0] >= '0' && e[0] <= '2') c->k2_filter = e[0] - '0'; }
  CREATE_TRY(dev_alloc(&c->d.corr, B * 2 * MPE_MAX_LEDS));
  CREATE_TRY(dev_alloc(&c->d.n_corr, B));
  CREATE_TRY(dev_alloc(&c->d.pose, B * 16));
  CREATE_TRY(dev_alloc(&c->d.cov, B * 36));
  CREATE_TRY(dev_alloc(&c->d.ok, B));
  CREATE_TRY(dev_alloc(&c->d.iters, B));
  CREATE_TRY(dev_alloc(&c->d.updated, B));
  CREATE_TRY(dev_alloc(&c->d.rois, B));
  CREATE_TRY(dev_alloc(&c->d.results, B));
  CREATE_TRY(dev_alloc(&c->d.check_sums, B * MPE_MAX_LEDS * 3));
  CREATE_TRY(dev_alloc(&c->d.check_cnt, B * 2));
  CREATE_TRY(dev_alloc(&c->d.hot_tiles, B * c->flags_per_frame));
  CREATE_TRY(dev_alloc(&c->d.pool, B * kPoolPerFrame));
  CREATE_TRY(dev_alloc(&c->d.counters, 2 * kLanes));
  CREATE_TRY(dev_alloc(&c->d.tile_list, B * c->flags_per_frame));
  CREATE_TRY(dev_alloc(&c->d.tile_count, 1));
  CREATE_TRY(dev_alloc(&c->d.k2_list, B));
  CREATE_TRY(dev_alloc(&c->d.k2_count, 1));
  CREATE_TRY(dev_alloc(&c->d.streams, B));
  CREATE_TRY(dev_alloc(&c->d.result_rois, B));
  CREATE_TRY(dev_alloc(&c->d.pred_px, B * MPE_MAX_LEDS * 2));
  CREATE_TRY(dev_alloc(&c->d.masks, 6 * B));
  CREATE_TRY(dev_alloc(&c->d.track_flags, B));
  CREATE_TRY(dev_alloc(&c->d.times, B));
  CREATE_TRY(launch_track_reset(c->d.streams, max_batch, c->own_stream));
  CREATE_TRY(cudaMemset(c->d.n_corr, 0, B * sizeof(int)));
  CREATE_TRY(cudaMemset(c->d.flags, 0, B * sizeof(int)));
  CREATE_TRY(cudaMemset(c->d.corr, 0, B * 2 * MPE_MAX_LEDS * sizeof(uint32_t)));
  CREATE_TRY(cudaMemset(c->d.rowflags, 0, B * c->flags_per_frame * sizeof(uint32_t)));
  CREATE_TRY(cudaMallocHost((void**)&c->h_results, B * sizeof(mpe_result)));
  CREATE_TRY(cudaMallocHost((void**)&c->h_times, B * sizeof(double)));
  {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, c->h_results, 0) == cudaSuccess) c->h_results_dev = (mpe_result*)dp; else cudaGetLastError();
    dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, c->h_times, 0) == cudaSuccess) c->h_times_dev = (const double*)dp; else cudaGetLastError();
  }
  for (int i = 0; i < 10; ++i) CREATE_TRY(cudaEventCreate(&c->ev[i]));
  // PoseEstimator::PoseEstimator() defaults (pose_estimator.cpp:36-39)
  c->pp.back_projection_pixel_tolerance = 3;
  c->pp.back_proj_sq_max = sqrt_less_than_bound(3.0);
  c->pp.nearest_neighbour_pixel_tolerance = 5;
  c->pp.certainty_threshold = 0.75;
  c->pp.valid_correspondence_threshold = 0.7;
  c->params.back_projection_pixel_tolerance = 3;
  c->params.nearest_neighbour_pixel_tolerance = 5;
  c->params.certainty_threshold = 0.75;
  c->params.valid_correspondence_threshold = 0.7;
#undef CREATE_TRY
  *out = c;
  return MPE_OK;
}

void mpe_destroy(mpe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->d.frames); cudaFree(c->d.rowflags); cudaFree(c->d.mask); cudaFree(c->d.n_det); cudaFree(c->d.flags);
  cudaFree(c->d.det); cudaFree(c->d.centers); cudaFree(c->d.hist); cudaFree(c->d.combos); cudaFree(c->d.triples); cudaFree(c->d.corr);
  cudaFree(c->d.n_corr); cudaFree(c->d.pose); cudaFree(c->d.cov); cudaFree(c->d.ok); cudaFree(c->d.iters);
  cudaFree(c->d.updated); cudaFree(c->d.rois); cudaFree(c->d.results); cudaFree(c->d.check_sums); cudaFree(c->d.check_cnt); cudaFree(c->d.hot_tiles); cudaFree(c->d.pool); cudaFree(c->d.counters); cudaFree(c->d.tile_list); cudaFree(c->d.tile_count); cudaFree(c->d.k2_list); cudaFree(c->d.k2_count); cudaFree(c->d.streams); cudaFree(c->d.result_rois);
  cudaFree(c->d.pred_px); cudaFree(c->d.masks); cudaFree(c->d.track_flags); cudaFree(c->d.times);
  if (c->h_results) cudaFreeHost(c->h_results);
  if (c->h_times) cudaFreeHost(c->h_times);
  for (int i = 0; i < 10; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (auto e : c->chunk_events) cudaEventDestroy(e);
  for (int gi = 0; gi < 2; ++gi) if (c->graph_exec[gi]) cudaGraphExecDestroy(c->graph_exec[gi]);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
}

const char* mpe_last_error(const mpe_ctx* c) { return c ? c->err.c_str() : "null context"; }

int mpe_set_stream(mpe_ctx* c, void* cuda_stream) {
  if (!c) return MPE_E_INVALID;
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return MPE_OK;
}

int mpe_set_camera(mpe_ctx* c, const double K[9], const double* D, int nD) {
  if (!c || !K) return MPE_E_INVALID;
  if (!(nD == 0 || nD == 4 || nD == 5 || nD == 8 || nD == 12)) return fail(c, MPE_E_UNSUPPORTED, "distortion vector must have 0, 4, 5, 8 or 12 coefficients");
  if (nD > 0 && !D) return MPE_E_INVALID;
  for (int i = 0; i < 9; ++i) c->cam.K[i] = K[i];
  for (int i = 0; i < MPE_MAX_DIST; ++i) c->cam.D[i] = (i < nD) ? D[i] : 0.0;
  c->cam.nD = nD;
  c->have_camera = true;
  ++c->cfg_version;
  return MPE_OK;
}

int mpe_set_markers(mpe_ctx* c, const double* xyz, int n) {
  if (!c || !xyz || n < 1 || n > MPE_MAX_LEDS) return fail(c, MPE_E_INVALID, "marker count must be in [1, MPE_MAX_LEDS]");
  c->pp.n_obj = n;
  for (int i = 0; i < 3 * n; ++i) c->pp.markers[i] = xyz[i];
  c->pp.histogram_threshold = (n >= 3) ? num_combinations_ref((unsigned)n, 3u) : 0u;   // pose_estimator.cpp:54
  // LED-triple table of the K2 sweep (world frame of every ordered marker triple), computed on the device with the kernels' arithmetic
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, launch_marker_triples(c->pp, c->d.triples, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  ++c->launches;
  c->have_markers = true;
  ++c->cfg_version;
  return MPE_OK;
}

int mpe_set_params(mpe_ctx* c, const mpe_params* p) {
  if (!c || !p) return MPE_E_INVALID;
  int radius;
  uint32_t taps[kMaxTaps];
  if (!gaussian_taps_8u(p->gaussian_sigma, &radius, taps))
    return fail(c, MPE_E_UNSUPPORTED, "gaussian_sigma must give a kernel radius in [1, " + std::to_string(kMaxRadius) + "] (sigma in about (0.09, 6.08))");
  c->params = *p;
  c->pp.back_projection_pixel_tolerance = p->back_projection_pixel_tolerance;
  c->pp.back_proj_sq_max = sqrt_less_than_bound(p->back_projection_pixel_tolerance);
  c->pp.nearest_neighbour_pixel_tolerance = p->nearest_neighbour_pixel_tolerance;
  c->pp.certainty_threshold = p->certainty_threshold;
  c->pp.valid_correspondence_threshold = p->valid_correspondence_threshold;
  c->have_params = true;
  ++c->cfg_version;
  return MPE_OK;
}

int mpe_set_histogram_threshold(mpe_ctx* c, uint32_t t) { if (!c) return MPE_E_INVALID; c->pp.histogram_threshold = t; ++c->cfg_version; return MPE_OK; }
uint32_t mpe_get_histogram_threshold(const mpe_ctx* c) { return c ? c->pp.histogram_threshold : 0u; }

int mpe_synchronize(mpe_ctx* c) {
  if (!c) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MPE_OK;
}

// ---------------------------------------------------------------------------------------- stage calls
int mpe_find_leds(mpe_ctx* c, const uint8_t* image, int pitch, int width, int height, mpe_rect roi, double* px_out,
                  float* centers_out, int* n_out, int* flags_out) {
  int rc = check_configured(c, false);
  if (rc != MPE_OK) return rc;
  if (!image || !n_out || width < 1 || height < 1 || pitch < width) return fail(c, MPE_E_INVALID, "bad image arguments");
  if (width > c->max_w || height > c->max_h) return fail(c, MPE_E_CAPACITY, "image larger than the context capacity");
  if (roi.x < 0 || roi.y < 0 || roi.width < 1 || roi.height < 1 || roi.x + roi.width > width || roi.y + roi.height > height)
    return fail(c, MPE_E_INVALID, "ROI outside the image (cv::Mat::operator() would throw)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CUDA_TRY(c, cudaMemcpy2DAsync(c->d.frames, c->pitch, image, pitch, width, height, cudaMemcpyHostToDevice, st));
  FrameSource src{c->d.frames, c->pitch, (long long)c->pitch * c->max_h, width, height, 1};
  Roi r{roi.x, roi.y, roi.width, roi.height};
  rc = run_find_leds(c, src, 0, 1, 0, r, nullptr, st);
  if (rc != MPE_OK) return rc;
  int n = 0, fl = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&n, c->d.n_det, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaMemcpyAsync(&fl, c->d.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (n > MPE_MAX_BLOBS) n = MPE_MAX_BLOBS;
  if (n > 0) {
    if (px_out) CUDA_TRY(c, cudaMemcpy(px_out, c->d.det, (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (centers_out) CUDA_TRY(c, cudaMemcpy(centers_out, c->d.centers, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost));
  }
  *n_out = n;
  if (flags_out) *flags_out = fl;
  return MPE_OK;
}

static int upload_detections(mpe_ctx* c, const double* det, int n_det, cudaStream_t st) {
  if (n_det < 0 || n_det > MPE_MAX_DET || (n_det > 0 && !det)) return fail(c, MPE_E_INVALID, "n_det must be in [0, MPE_MAX_DET]");
  if (n_det > 0) CUDA_TRY(c, cudaMemcpyAsync(c->d.det, det, (size_t)n_det * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->d.n_det, &n_det, sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemsetAsync(c->d.flags, 0, sizeof(int), st));
  return MPE_OK;
}

static int upload_correspondences(mpe_ctx* c, const uint32_t* corr, int k, cudaStream_t st) {
  if (k < 0 || k > MPE_MAX_LEDS || (k > 0 && !corr)) return fail(c, MPE_E_INVALID, "k must be in [0, MPE_MAX_LEDS]");
  for (int i = 0; i < k; ++i) {
    if (corr[2 * i] < 1 || corr[2 * i] > (uint32_t)c->pp.n_obj) return fail(c, MPE_E_INVALID, "correspondence LED index out of range");
  }
  if (k > 0) CUDA_TRY(c, cudaMemcpyAsync(c->d.corr, corr, (size_t)k * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->d.n_corr, &k, sizeof(int), cudaMemcpyHostToDevice, st));
  return MPE_OK;
}

int mpe_initialise(mpe_ctx* c, const double* det, int n_det, uint32_t* hist_out, uint32_t* corr_out, int* k_out, double pose_out[16], int* ok) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  rc = upload_detections(c, det, n_det, st);
  if (rc != MPE_OK) return rc;
  int zero = 0;
  CUDA_TRY(c, cudaMemcpyAsync(c->d.n_corr, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  rc = run_sweep(c, 0, 1, st, nullptr);
  if (rc != MPE_OK) return rc;
  CUDA_TRY(c, cudaMemsetAsync(c->d.pose, 0, 16 * sizeof(double), st));
  rc = run_refine(c, 0, 1, 1, st, nullptr);
  if (rc != MPE_OK) return rc;
  int k = 0, okv = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&k, c->d.n_corr, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaMemcpyAsync(&okv, c->d.ok, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (hist_out && n_det > 0) CUDA_TRY(c, cudaMemcpy(hist_out, c->d.hist, (size_t)n_det * c->pp.n_obj * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (corr_out && k > 0) CUDA_TRY(c, cudaMemcpy(corr_out, c->d.corr, (size_t)k * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (pose_out) CUDA_TRY(c, cudaMemcpy(pose_out, c->d.pose, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  if (k_out) *k_out = k;
  if (ok) *ok = okv;
  return MPE_OK;
}

int mpe_check_correspondences(mpe_ctx* c, const double* det, int n_det, const uint32_t* corr, int k, double pose_out[16], int* ok) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  rc = upload_detections(c, det, n_det, st);
  if (rc != MPE_OK) return rc;
  rc = upload_correspondences(c, corr, k, st);
  if (rc != MPE_OK) return rc;
  for (int i = 0; i < k; ++i)
    if (corr[2 * i + 1] < 1 || corr[2 * i + 1] > (uint32_t)n_det) return fail(c, MPE_E_INVALID, "checkCorrespondences needs detection indices in [1, n_det]");
  CUDA_TRY(c, cudaMemsetAsync(c->d.pose, 0, 16 * sizeof(double), st));
  rc = run_refine(c, 0, 1, 1, st, nullptr);
  if (rc != MPE_OK) return rc;
  int okv = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&okv, c->d.ok, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (pose_out) CUDA_TRY(c, cudaMemcpy(pose_out, c->d.pose, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  if (ok) *ok = okv;
  return MPE_OK;
}

int mpe_optimise_pose(mpe_ctx* c, const double* det, int n_det, const uint32_t* corr, int k, double pose_io[16], double cov_out[36], int* iters_out) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  if (!pose_io) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  rc = upload_detections(c, det, n_det, st);
  if (rc != MPE_OK) return rc;
  rc = upload_correspondences(c, corr, k, st);
  if (rc != MPE_OK) return rc;
  for (int i = 0; i < k; ++i)
    if (corr[2 * i + 1] > (uint32_t)n_det) return fail(c, MPE_E_INVALID, "detection index out of range");
  CUDA_TRY(c, cudaMemcpyAsync(c->d.pose, pose_io, 16 * sizeof(double), cudaMemcpyHostToDevice, st));
  rc = run_refine(c, 0, 1, 2, st, nullptr);
  if (rc != MPE_OK) return rc;
  int it = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&it, c->d.iters, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  CUDA_TRY(c, cudaMemcpy(pose_io, c->d.pose, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  if (cov_out) CUDA_TRY(c, cudaMemcpy(cov_out, c->d.cov, 36 * sizeof(double), cudaMemcpyDeviceToHost));
  if (iters_out) *iters_out = it;
  return MPE_OK;
}

int mpe_p3p_compute_poses(mpe_ctx* c, const double* feature_vectors, const double* world_points, int n, double* solutions, int* status) {
  if (!c || n < 0 || (n > 0 && (!feature_vectors || !world_points || !solutions || !status))) return MPE_E_INVALID;
  if (n == 0) return MPE_OK;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  // one allocation for the four arrays, released on every path (doubles first: alignment)
  struct Scratch { void* p = nullptr; ~Scratch() { if (p) cudaFree(p); } } scratch;
  const size_t nf = (size_t)n * 9, ns = (size_t)n * 48;
  CUDA_TRY(c, cudaMalloc(&scratch.p, (2 * nf + ns) * sizeof(double) + (size_t)n * sizeof(int)));
  double* df = (double*)scratch.p;
  double* dP = df + nf;
  double* ds = dP + nf;
  int* dst = (int*)(ds + ns);
  CUDA_TRY(c, cudaMemcpyAsync(df, feature_vectors, nf * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(dP, world_points, nf * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, launch_p3p_batch(df, dP, n, ds, dst, st));
  ++c->launches;
  CUDA_TRY(c, cudaMemcpyAsync(solutions, ds, ns * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaMemcpyAsync(status, dst, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  return MPE_OK;
}

// ---------------------------------------------------------------------------------------- batch drivers
int mpe_estimate_batch_device_async(mpe_ctx* c, const uint8_t* frames_device, int pitch, long long frame_stride, int width, int height, int n_frames) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  if (!frames_device || n_frames < 1) return MPE_E_INVALID;
  if (n_frames > c->max_batch || width > c->max_w || height > c->max_h) return fail(c, MPE_E_CAPACITY, "batch or image larger than the context capacity");
  CUDA_TRY(c, cudaSetDevice(c->device));
  FrameSource src{frames_device, pitch, frame_stride, width, height, n_frames};
  return run_cold_pipelined(c, src, n_frames, c->stream);
}

int mpe_fetch_results(mpe_ctx* c, int n_frames, mpe_result* results) {
  if (!c || !results || n_frames < 1 || n_frames > c->max_batch) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpyAsync(c->h_results, c->d.results, (size_t)n_frames * sizeof(mpe_result), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  std::memcpy(results, c->h_results, (size_t)n_frames * sizeof(mpe_result));
  return MPE_OK;
}

int mpe_estimate_batch_device(mpe_ctx* c, const uint8_t* frames_device, int pitch, long long frame_stride, int width, int height, int n_frames, mpe_result* results) {
  int rc = mpe_estimate_batch_device_async(c, frames_device, pitch, frame_stride, width, height, n_frames);
  if (rc != MPE_OK) return rc;
  return mpe_fetch_results(c, n_frames, results);
}

int mpe_estimate_batch(mpe_ctx* c, const uint8_t* frames, int pitch, long long frame_stride, int width, int height, int n_frames, mpe_result* results) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  if (!frames || !results || n_frames < 1 || pitch < width) return MPE_E_INVALID;
  if (width > c->max_w || height > c->max_h) return fail(c, MPE_E_CAPACITY, "image larger than the context capacity");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const long long dev_stride = (long long)c->pitch * c->max_h;
  FrameSource src{c->d.frames, c->pitch, dev_stride, width, height, c->max_batch};
  // chunked pipeline: H2D of chunk i+1 (copy stream) overlaps the kernels of chunk i (compute stream)
  int chunk = 128;
  if (chunk > c->max_batch) chunk = c->max_batch;
  for (int super = 0; super < n_frames; super += c->max_batch) {
    int n_super = n_frames - super < c->max_batch ? n_frames - super : c->max_batch;
    int n_chunks = (n_super + chunk - 1) / chunk;
    while ((int)c->chunk_events.size() < n_chunks) {
      cudaEvent_t e;
      CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->chunk_events.push_back(e);
    }
    for (int ci = 0; ci < n_chunks; ++ci) {
      int f0 = ci * chunk;
      int n = n_super - f0 < chunk ? n_super - f0 : chunk;
      const uint8_t* hsrc = frames + (size_t)(super + f0) * frame_stride;
      uint8_t* ddst = c->d.frames + (size_t)f0 * dev_stride;
      if (pitch == c->pitch && frame_stride == dev_stride) {
        CUDA_TRY(c, cudaMemcpyAsync(ddst, hsrc, (size_t)n * dev_stride, cudaMemcpyHostToDevice, c->copy_stream));
      } else if (frame_stride == (long long)pitch * height && c->max_h == height) {
        // frames are contiguous: one 2-D copy of n*height rows
        CUDA_TRY(c, cudaMemcpy2DAsync(ddst, c->pitch, hsrc, pitch, width, (size_t)n * height, cudaMemcpyHostToDevice, c->copy_stream));
      } else {
        for (int f = 0; f < n; ++f)
          CUDA_TRY(c, cudaMemcpy2DAsync(ddst + (size_t)f * dev_stride, c->pitch, hsrc + (size_t)f * frame_stride, pitch, width, height,
                                        cudaMemcpyHostToDevice, c->copy_stream));
      }
      CUDA_TRY(c, cudaEventRecord(c->chunk_events[ci], c->copy_stream));
      CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->chunk_events[ci], 0));
      rc = run_cold(c, src, f0, n, f0, c->stream);
      if (rc != MPE_OK) return rc;
      CUDA_TRY(c, cudaMemcpyAsync(c->h_results + f0, c->d.results + f0, (size_t)n * sizeof(mpe_result), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    std::memcpy(results + super, c->h_results, (size_t)n_super * sizeof(mpe_result));
  }
  return MPE_OK;
}

int mpe_copy_poses_device(mpe_ctx* c, int n_frames, double* poses_device) {
  if (!c || !poses_device || n_frames < 1 || n_frames > c->max_batch) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpyAsync(poses_device, c->d.pose, (size_t)n_frames * 16 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return MPE_OK;
}

// ---- host-side helpers of the stage-by-stage mirrors (no context, no GPU): the same functions K4 runs per stream ------------
namespace {
mpe::DevCamera host_camera(const double K[9], const double* D, int nD) {
  mpe::DevCamera cam; std::memset(&cam, 0, sizeof(cam));
  for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
  for (int i = 0; i < nD && i < MPE_MAX_DIST; ++i) cam.D[i] = D[i];
  cam.nD = nD;
  return cam;
}
mpe::M4 m4_of(const double* p) { mpe::M4 m; std::memcpy(m.m, p, sizeof(m.m)); return m; }
}  // namespace

int mpe_host_predict_pose(const double previous_pose[16], const double current_pose[16], double previous_time, double current_time,
                          double time_to_predict, double predicted_pose_out[16]) {
  if (!previous_pose || !current_pose || !predicted_pose_out) return MPE_E_INVALID;
  const mpe::M4 r = mpe::predict_pose(m4_of(previous_pose), m4_of(current_pose), previous_time, current_time, time_to_predict);
  std::memcpy(predicted_pose_out, r.m, sizeof(r.m));
  return MPE_OK;
}
int mpe_host_project_markers(const double K[9], const double pose[16], const double* markers_xyz, int n, double* pixels_out) {
  if (!K || !pose || !markers_xyz || !pixels_out || n < 0) return MPE_E_INVALID;
  const mpe::M4 T = m4_of(pose);
  for (int i = 0; i < n; ++i) mpe::project2d(K, T, markers_xyz[3 * i], markers_xyz[3 * i + 1], markers_xyz[3 * i + 2], pixels_out[2 * i], pixels_out[2 * i + 1]);
  return MPE_OK;
}
int mpe_host_determine_roi(const double* pixels, int n, int width, int height, int border, const double K[9], const double* D, int nD, mpe_rect* roi_out) {
  if (!pixels || !K || !roi_out || n < 0 || nD < 0 || (nD > 0 && !D)) return MPE_E_INVALID;
  const mpe::Roi r = mpe::determine_roi(host_camera(K, D, nD), pixels, n, width, height, border);
  roi_out->x = r.x; roi_out->y = r.y; roi_out->width = r.w; roi_out->height = r.h;
  return MPE_OK;
}
int mpe_host_exponential_map(const double twist[6], double pose_out[16]) {
  if (!twist || !pose_out) return MPE_E_INVALID;
  const mpe::M4 r = mpe::exponential_map(twist);
  std::memcpy(pose_out, r.m, sizeof(r.m));
  return MPE_OK;
}
int mpe_host_logarithm_map(const double pose[16], double twist_out[6]) {
  if (!pose || !twist_out) return MPE_E_INVALID;
  mpe::logarithm_map(m4_of(pose), twist_out);
  return MPE_OK;
}

// host-only debug export: the fixed-point Gaussian taps mpe_set_params derives from sigma (no context, no GPU)
int mpe_debug_gaussian_taps(double sigma, int* radius_out, uint32_t* taps_out, int taps_capacity) {
  if (!radius_out || !taps_out) return MPE_E_INVALID;
  uint32_t taps[kMaxTaps];
  int r = 0;
  if (!gaussian_taps_8u(sigma, &r, taps)) return MPE_E_INVALID;
  if (2 * r + 1 > taps_capacity) return MPE_E_CAPACITY;
  for (int i = 0; i < 2 * r + 1; ++i) taps_out[i] = taps[i];
  *radius_out = r;
  return MPE_OK;
}

int mpe_copy_results_device(mpe_ctx* c, int n_frames, mpe_result* results_device) {
  if (!c || !results_device || n_frames < 1 || n_frames > c->max_batch) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpyAsync(results_device, c->d.results, (size_t)n_frames * sizeof(mpe_result), cudaMemcpyDeviceToDevice, c->stream));
  return MPE_OK;
}

// FP64 peak probe: kProbeChains independent DFMA chains per thread, enough CTAs to fill every SM.
namespace {
constexpr int kProbeChains = 8, kProbeIters = 4096;
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, double a, double b) {
  double x[kProbeChains];
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) x[i] = a + (double)(threadIdx.x + i);
  for (int it = 0; it < kProbeIters; ++it) {
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i) x[i] = fma(x[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;          // never true: keeps the chains alive
}
}  // namespace

int mpe_probe_fp64_peak(mpe_ctx* c, double* tflops_out) {
  if (!c || !tflops_out) return MPE_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  double* d = nullptr;
  CUDA_TRY(c, cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = c->n_sms * 8;
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, c->stream);
    fp64_probe_kernel<<<grid, 256, 0, c->stream>>>(d, 0.5, 0.999999);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1); return fail(c, MPE_E_CUDA, std::string("fp64 probe: ") + cudaGetErrorString(e)); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * (double)grid * 256.0 * kProbeChains * kProbeIters;
    if (rep > 0 && ms > 0) best = std::fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops_out = best;
  return MPE_OK;
}

int mpe_streams_reset(mpe_ctx* c, int n_streams) {
  if (!c || n_streams < 1 || n_streams > c->max_batch) return fail(c, MPE_E_INVALID, "n_streams must be in [1, max_batch]");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, launch_track_reset(c->d.streams, n_streams, c->stream));
  ++c->launches;
  c->stream_tracking.assign(c->max_batch, 0);
  c->n_tracking = 0;
  return MPE_OK;
}

int mpe_streams_set_frame_map(mpe_ctx* c, const int* frame_index_device, int n_frames_in_buffer) {
  if (!c) return MPE_E_INVALID;
  c->frame_map = frame_index_device;
  c->frame_map_total = n_frames_in_buffer;
  return MPE_OK;
}

// Enqueues one estimateBodyPose step (pose_estimator.cpp:62-147) for n streams on `st`; the time stamps are already in c->d.times.
constexpr int kPdlMaxStreams = 64;
static int enqueue_streams_step(mpe_ctx* c, const FrameSource& src, int n, cudaStream_t st, bool fast, bool mapped_io) {
  const int width = src.width, height = src.height;
  const size_t B = (size_t)c->max_batch;
  TrackArgs t{};
  t.n = n; t.img_w = width; t.img_h = height; t.roi_border = c->params.roi_border_thickness;
  t.state = c->d.streams; t.times = mapped_io ? c->h_times_dev : c->d.times; t.cam = c->cam; t.pp = c->pp;
  t.rois = c->d.rois; t.result_rois = c->d.result_rois; t.pred_px = c->d.pred_px;
  t.mode = c->d.masks; t.done = c->d.masks + B; t.a_retry = c->d.masks + 2 * B; t.a_check = c->d.masks + 3 * B;
  t.a_init = c->d.masks + 4 * B; t.a_gn = c->d.masks + 5 * B;
  t.track_flags = c->d.track_flags;
  t.fast = fast ? 1 : 0;
  t.n_det = c->d.n_det; t.flags = c->d.flags; t.det = c->d.det; t.centers = c->d.centers;
  t.corr = c->d.corr; t.n_corr = c->d.n_corr; t.pose_io = c->d.pose; t.cov = c->d.cov; t.ok = c->d.ok; t.iters = c->d.iters; t.updated = c->d.updated;
  Roi full{0, 0, width, height};
  int rc;
  // few cameras: programmatic dependent launches hide the launch latency between the step's small kernels (mpe_internal.cuh)
  static int pdl_env = -1;
  if (pdl_env < 0) { const char* e = getenv("MPE_PDL"); pdl_env = e ? atoi(e) : 1; }
  PdlScope pdl(pdl_env != 0 && n <= kPdlMaxStreams);
  g_trace.begin(st);
  CUDA_TRY(c, launch_track_begin(t, st));                                                        // predictWithROI
  g_trace.mark("predict", st);
  rc = run_find_leds(c, src, 0, n, 0, full, c->d.rois, st, kTrackTileWidthPx, c->frame_map, nullptr);   // findLeds(ROI)
  if (rc != MPE_OK) return rc;
  g_trace.mark("findLeds", st);
  CUDA_TRY(c, launch_track_after_detect(t, 0, st));                                              // + empties the ROI of streams that do not retry
  g_trace.mark("match", st);
  if (!fast) {
    rc = run_find_leds(c, src, 0, n, 0, full, c->d.rois, st, kTrackTileWidthPx, c->frame_map, t.a_retry);  // whole-image retry (only where needed)
    if (rc != MPE_OK) return rc;
    CUDA_TRY(c, launch_track_after_detect(t, 1, st));
    ++c->launches;
  }
  rc = run_refine(c, 0, n, 1, st, t.a_check, t.a_gn, t.a_init);                                  // checkCorrespondences on the NN matches; ok -> GN, else -> initialise()
  if (rc != MPE_OK) return rc;
  g_trace.mark("check+kabsch", st);
  if (!fast) {
    rc = run_sweep(c, 0, n, st, t.a_init);                                                       // initialise(): cold streams + failed checks
    if (rc != MPE_OK) return rc;
    rc = run_refine(c, 0, n, 1, st, t.a_init, t.a_gn, nullptr);                                  // its check; ok -> GN
    if (rc != MPE_OK) return rc;
  }
  rc = run_refine(c, 0, n, 2, st, t.a_gn);                                                       // optimisePose
  if (rc != MPE_OK) return rc;
  g_trace.mark("gauss-newton", st);
  CUDA_TRY(c, launch_track_finish(t, c->d.results, mapped_io ? c->h_results_dev : nullptr, st));
  g_trace.mark("finish", st);
  c->launches += 3;
  return MPE_OK;
}

// The step as one CUDA-graph replay: captured once per (frame buffer, geometry, configuration) key, then a single
// cudaGraphLaunch replaces ~30 kernel/memset launches — what makes a one-camera, one-image-at-a-time caller (the way MPENode
// drives the reference, monocular_pose_estimator.cpp:133-159) latency-competitive.  Per-kernel event timing disables replay.
static int run_streams_step(mpe_ctx* c, const FrameSource& src, int n, const double* times, bool fetch, bool fast = false) {
  const int gi = fast ? 1 : 0;
  cudaStream_t st = c->stream;
  // A few cameras whose records the caller waits for: the time stamps are read, and the records written, in place in page-locked
  // host memory — no copy operations around the step (the step is synchronised before the call returns, see finish_step).
  const bool mapped_io = fetch && n <= kPdlMaxStreams && c->mapped_io && c->h_times_dev && c->h_results_dev;
  if (mapped_io) std::memcpy(c->h_times, times, (size_t)n * sizeof(double));
  else CUDA_TRY(c, cudaMemcpyAsync(c->d.times, times, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  const bool copy_back = fetch && !mapped_io;
  if (!c->use_graphs || c->timing) {
    int rc = enqueue_streams_step(c, src, n, st, fast, mapped_io);
    if (rc != MPE_OK) return rc;
    if (copy_back) CUDA_TRY(c, cudaMemcpyAsync(c->h_results, c->d.results, (size_t)n * sizeof(mpe_result), cudaMemcpyDeviceToHost, st));
    return MPE_OK;
  }
  mpe_ctx::StepGraphKey key{src.base, src.pitch, src.frame_stride, src.width, src.height, n, c->frame_map, c->frame_map_total, (fetch ? 1 : 0) | (mapped_io ? 2 : 0),
                            c->cfg_version, st};
  if (!c->graph_exec[gi] || !(key == c->graph_key[gi])) {
    if (!c->have_pending[gi] || !(key == c->pending_key[gi])) {      // first use of this key: plain launches
      c->pending_key[gi] = key; c->have_pending[gi] = true;
      int rc = enqueue_streams_step(c, src, n, st, fast, mapped_io);
      if (rc != MPE_OK) return rc;
      if (copy_back) CUDA_TRY(c, cudaMemcpyAsync(c->h_results, c->d.results, (size_t)n * sizeof(mpe_result), cudaMemcpyDeviceToHost, st));
      return MPE_OK;
    }
    if (c->graph_exec[gi]) { cudaGraphExecDestroy(c->graph_exec[gi]); c->graph_exec[gi] = nullptr; }
    const long long l0 = c->launches;
    CUDA_TRY(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_streams_step(c, src, n, st, fast, mapped_io);
    cudaError_t ce = cudaSuccess;
    if (rc == MPE_OK && copy_back) ce = cudaMemcpyAsync(c->h_results, c->d.results, (size_t)n * sizeof(mpe_result), cudaMemcpyDeviceToHost, st);
    cudaGraph_t g = nullptr;
    cudaError_t ee = cudaStreamEndCapture(st, &g);
    if (rc != MPE_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess || ee != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      return fail(c, MPE_E_CUDA, std::string("stream capture of the tracking step failed: ") + cudaGetErrorString(ce != cudaSuccess ? ce : ee));
    }
    cudaError_t ie = cudaGraphInstantiate(&c->graph_exec[gi], g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) { c->graph_exec[gi] = nullptr; return fail(c, MPE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
    c->graph_launches[gi] = c->launches - l0;
    c->launches = l0;                        // nothing ran during the capture
    c->graph_key[gi] = key;
  }
  CUDA_TRY(c, cudaGraphLaunch(c->graph_exec[gi], st));
  c->launches += c->graph_launches[gi];
  ++c->graph_replays;
  return MPE_OK;
}

static int finish_step(mpe_ctx* c, int n, mpe_result* results) {
  if (!results) return MPE_OK;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  std::memcpy(results, c->h_results, (size_t)n * sizeof(mpe_result));
  // host mirror of "this stream is tracking" (it_since_initialized_ never returns to 0, pose_estimator.cpp:806-809)
  if ((int)c->stream_tracking.size() < c->max_batch) c->stream_tracking.assign(c->max_batch, 0);
  for (int i = 0; i < n; ++i)
    if (results[i].updated && !c->stream_tracking[i]) { c->stream_tracking[i] = 1; ++c->n_tracking; }
  return MPE_OK;
}

// The short step serves the latency case: a handful of cameras, all of them tracking.  It leaves out the stages that a tracking
// stream almost never needs (whole-image retry: 5 launches; brute-force re-initialisation and its check: 6 launches); a stream that
// does need one is left untouched and flagged, and the complete step is run right after (the records are read on the host anyway).
constexpr int kShortStepMaxStreams = 64;
static bool want_short_step(const mpe_ctx* c, int n, bool fetch) {
  return c->short_steps && fetch && n <= kShortStepMaxStreams && c->n_tracking >= n && !c->timing;
}
static bool short_step_incomplete(const mpe_ctx* c, int n) {
  for (int i = 0; i < n; ++i) if (c->h_results[i].flags & mpe::kFlagNeedsFullStep) return true;
  return false;
}

static int run_short_then_full(mpe_ctx* c, const FrameSource& src, int n, const double* times) {
  const bool fast = want_short_step(c, n, true);
  int rc = run_streams_step(c, src, n, times, true, fast);
  if (rc != MPE_OK || !fast) return rc;
  ++c->short_step_count;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (!short_step_incomplete(c, n)) return MPE_OK;
  ++c->short_step_fallbacks;
  return run_streams_step(c, src, n, times, true, false);
}

// Is `p` page-locked host memory the GPU can read in place (cudaHostAlloc / cudaHostRegister)?  Returns its device alias.
static const uint8_t* device_alias_of_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return (const uint8_t*)at.devicePointer;
}

// One estimateBodyPose step for n_streams independent PoseEstimators (pose_estimator.cpp:62-147) without a host round trip.
int mpe_streams_step_device(mpe_ctx* c, const uint8_t* frames_device, int pitch, long long frame_stride, int width, int height,
                            int n_streams, const double* times, mpe_result* results) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  if (!frames_device || !times || n_streams < 1) return MPE_E_INVALID;
  if (n_streams > c->max_batch || width > c->max_w || height > c->max_h) return fail(c, MPE_E_CAPACITY, "streams or image larger than the context capacity");
  if (c->cam.nD < 5) return fail(c, MPE_E_UNSUPPORTED, "LEDDetector::distortPoints reads five distortion coefficients (led_detector.cpp:190-194)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int total = c->frame_map ? c->frame_map_total : n_streams;
  FrameSource src{frames_device, pitch, frame_stride, width, height, total};
  const bool fast = want_short_step(c, n_streams, results != nullptr);
  rc = run_streams_step(c, src, n_streams, times, results != nullptr, fast);
  if (rc != MPE_OK) return rc;
  if (fast) {
    ++c->short_step_count;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (short_step_incomplete(c, n_streams)) {
      ++c->short_step_fallbacks;
      rc = run_streams_step(c, src, n_streams, times, true, false);
      if (rc != MPE_OK) return rc;
    }
  }
  return finish_step(c, n_streams, results);
}

// Host-image variant: the per-image call of a camera driver (MPENode::imageCallback -> estimateBodyPose,
// monocular_pose_estimator.cpp:133-159).  Image s (HOST memory, pitch bytes per row) belongs to stream s.
int mpe_streams_step(mpe_ctx* c, const uint8_t* frames, int pitch, long long frame_stride, int width, int height, int n_streams,
                     const double* times, mpe_result* results) {
  int rc = check_configured(c, true);
  if (rc != MPE_OK) return rc;
  if (!frames || !times || !results || n_streams < 1 || pitch < width) return MPE_E_INVALID;
  if (n_streams > c->max_batch || width > c->max_w || height > c->max_h) return fail(c, MPE_E_CAPACITY, "streams or image larger than the context capacity");
  if (c->cam.nD < 5) return fail(c, MPE_E_UNSUPPORTED, "LEDDetector::distortPoints reads five distortion coefficients (led_detector.cpp:190-194)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const long long dev_stride = (long long)c->pitch * c->max_h;
  // ---- ingest: copy the whole images, or let the kernels read the ROI tiles in place over PCIe (zero copy)
  const uint8_t* alias = nullptr;
  if (c->ingest_mode != MPE_INGEST_COPY && (pitch % 16) == 0 && (frame_stride % 16) == 0) {
    alias = device_alias_of_pinned(frames);
    if (alias && ((uintptr_t)alias & 15)) alias = nullptr;
  }
  if (c->ingest_mode == MPE_INGEST_ZERO_COPY && !alias)
    return fail(c, MPE_E_INVALID, "zero-copy ingest needs page-locked images (cudaHostAlloc / cudaHostRegister), 16-byte aligned, pitch and frame stride multiples of 16");
  // AUTO: in place only when every stream is tracking (ROI search: ~1/7 of the image bytes); whole-image searches are
  // cheaper through one bulk copy (53 GB/s against ~24 GB/s for tile reads over PCIe)
  const bool zero_copy = alias && (c->ingest_mode == MPE_INGEST_ZERO_COPY || c->n_tracking >= n_streams);
  const int* saved_map = c->frame_map;
  const int saved_total = c->frame_map_total;
  c->frame_map = nullptr; c->frame_map_total = 0;          // host images are one per stream
  if (zero_copy) {
    ++c->zero_copy_steps;
    FrameSource src{alias, pitch, frame_stride, width, height, n_streams};
    rc = run_short_then_full(c, src, n_streams, times);
  } else {
    ++c->copy_steps;
    c->h2d_bytes_copied += (long long)n_streams * width * height;
    if (pitch == c->pitch && frame_stride == dev_stride) {
      CUDA_TRY(c, cudaMemcpyAsync(c->d.frames, frames, (size_t)n_streams * dev_stride, cudaMemcpyHostToDevice, st));
    } else if (frame_stride == (long long)pitch * height && c->max_h == height) {
      CUDA_TRY(c, cudaMemcpy2DAsync(c->d.frames, c->pitch, frames, pitch, width, (size_t)n_streams * height, cudaMemcpyHostToDevice, st));
    } else {
      for (int f = 0; f < n_streams; ++f)
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d.frames + (size_t)f * dev_stride, c->pitch, frames + (size_t)f * frame_stride, pitch, width, height,
                                      cudaMemcpyHostToDevice, st));
    }
    FrameSource src{c->d.frames, c->pitch, dev_stride, width, height, n_streams};
    rc = run_short_then_full(c, src, n_streams, times);
  }
  c->frame_map = saved_map; c->frame_map_total = saved_total;
  if (rc != MPE_OK) return rc;
  return finish_step(c, n_streams, results);
}

int mpe_set_ingest_mode(mpe_ctx* c, int mode) {
  if (!c || mode < MPE_INGEST_COPY || mode > MPE_INGEST_AUTO) return MPE_E_INVALID;
  c->ingest_mode = mode;
  return MPE_OK;
}

int mpe_get_ingest_stats(const mpe_ctx* c, long long* copy_steps, long long* zero_copy_steps, long long* h2d_bytes_copied) {
  if (!c) return MPE_E_INVALID;
  if (copy_steps) *copy_steps = c->copy_steps;
  if (zero_copy_steps) *zero_copy_steps = c->zero_copy_steps;
  if (h2d_bytes_copied) *h2d_bytes_copied = c->h2d_bytes_copied;
  return MPE_OK;
}

int mpe_set_k2_filter(mpe_ctx* c, int mode) { if (!c) return MPE_E_INVALID; c->k2_filter = (mode == 0) ? 0 : (mode == 1 ? 1 : 2); ++c->cfg_version; return MPE_OK; }

int mpe_set_graph_replay(mpe_ctx* c, int on) { if (!c) return MPE_E_INVALID; c->use_graphs = on != 0; return MPE_OK; }

int mpe_enable_kernel_timing(mpe_ctx* c, int on) { if (!c) return MPE_E_INVALID; c->timing = on != 0; return MPE_OK; }

int mpe_get_kernel_times(mpe_ctx* c, float ms_out[5]) {
  if (!c || !ms_out) return MPE_E_INVALID;
  if (!c->timing || !c->timing_pending) return fail(c, MPE_E_INVALID, "kernel timing not enabled or no batch run yet");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (int k = 0; k < 5; ++k) {
    CUDA_TRY(c, cudaEventSynchronize(c->ev[2 * k + 1]));
    CUDA_TRY(c, cudaEventElapsedTime(&ms_out[k], c->ev[2 * k], c->ev[2 * k + 1]));
  }
  return MPE_OK;
}

long long mpe_kernel_launch_count(const mpe_ctx* c) { return c ? c->launches : 0; }

// Host-only: the fields MPENode::imageCallback fills after estimateBodyPose (monocular_pose_estimator.cpp:160-190).  The quaternion
// follows the published algorithm of Eigen's Quaterniond(Matrix3d) constructor (trace branch, else the largest diagonal entry).
void mpe_pose_to_message(const double pose[16], const double cov[36], double position[3], double orientation_xyzw[4], double covariance[36]) {
  if (!pose) return;
  if (position) { position[0] = pose[3]; position[1] = pose[7]; position[2] = pose[11]; }
  if (orientation_xyzw) {
    auto m = [&](int r, int c) { return pose[4 * r + c]; };
    double q[4];                                             // x, y, z, w
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      q[3] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (m(2, 1) - m(1, 2)) * t;
      q[1] = (m(0, 2) - m(2, 0)) * t;
      q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
      q[i] = 0.5 * t;
      t = 0.5 / t;
      q[3] = (m(k, j) - m(j, k)) * t;
      q[j] = (m(j, i) + m(i, j)) * t;
      q[k] = (m(k, i) + m(i, k)) * t;
    }
    for (int e = 0; e < 4; ++e) orientation_xyzw[e] = q[e];
  }
  if (covariance && cov)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) covariance[j + 6 * i] = cov[6 * i + j];     // elems[j + 6*i] = cov(i, j)
}

}  // extern "C"
