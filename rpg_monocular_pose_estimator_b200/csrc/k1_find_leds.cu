// K1 — LEDDetector::findLeds on the GPU (reference: monocular_pose_estimator_lib/src/led_detector.cpp:35-112).
//
//   scan_kernel          (K1a) the streaming pass.  HBM-bound: every ROI byte is read exactly once from DRAM through
//                             TMA (cp.async.bulk.tensor) into a shared-memory ring and only asked "does this word hold
//                             a byte above the threshold"; the few hot words go to a pool, nothing else is written.
//                             Whole-image batches walk the frames x strips x column-tiles grid, per-frame ROIs (tracking)
//                             a tile work list built by build_tile_list_kernel.
//   blur_kernel          (K1c) exact THRESH_TOZERO + 8-bit fixed-point Gaussian + "blurred != 0" on the neighbourhood of
//                             the hot words only; writes a sparse 1-bit mask (rows that contain foreground) and one
//                             flag word per tile.
//   extract_blobs_kernel (K1b) external contours (cv::findContours RETR_EXTERNAL / CHAIN_APPROX_NONE),
//                             contourArea / boundingRect / moments, the four shape filters and
//                             cv::undistortPoints, one warp per frame working on the sparse mask.
//   (The first version fused K1a and K1c; see DESIGN.md section 4, decision 2, for why they are separate kernels.)
//
// Bit-exact contract (pinned against cv2 4.13):
//   blur    taps = getGaussianKernelBitExact -> 8.8 fixed point (sum 256), horizontal pass 8.8, vertical
//           pass 16.16, out = (acc + 32768) >> 16, BORDER_REFLECT_101 at the ROI edge.
//   contour Suzuki border following as OpenCV's icvFetchContour (start at the component's raster-first
//           pixel, clockwise search from NW for the closing point, counter-clockwise search afterwards),
//           components enclosed by another component are not reported, output in reverse raster order of
//           the start pixels.
//
// Why the blur is evaluated sparsely: a blurred pixel can only be non-zero if a thresholded pixel within
// the (2R+1)^2 window is non-zero.  The dense phase therefore only asks "does this 16-byte unit contain a
// byte > threshold" (about 0.4 instructions per pixel, SWAR on OR-ed words) — at 23 B/clk/SM of HBM
// bandwidth a dense 25-tap fixed-point blur (>10 instructions per pixel) would be issue-bound at ~1/3 of
// the memory roofline.  Only words whose neighbourhood is "hot" run the exact fixed-point arithmetic, and
// its result is identical to the dense evaluation.
#include "mpe_internal.cuh"
#include <cstdlib>

namespace mpe {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers (TMA + mbarrier)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 3-D tiled TMA load: tensor (x: u32 elements along a row, y: image row, z: frame) -> dense smem box.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
// cv::BORDER_REFLECT_101 index into [0, n)
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
  }
  return p;
}

// "does any byte of w exceed the threshold" — exact SWAR test.
//   T < 128 : byte > T  <=>  bit7(byte) | bit7((byte & 0x7f) + (127 - T))
//   T >= 128: byte > T  <=>  bit7(byte) & bit7((byte & 0x7f) + (255 - T))
template <bool kLow>
__device__ __forceinline__ uint32_t any_byte_gt(uint32_t w, uint32_t k) {
  uint32_t x = (w & 0x7f7f7f7fu) + k;
  return (kLow ? (x | w) : (x & w)) & 0x80808080u;
}

// TMA requires the innermost start coordinate to be a multiple of 16 bytes (measured on B200: any other value raises
// 'illegal instruction', tests/probes/tma_probe.cu), so tiles start at a 16-pixel boundary: first u32 element of the
// 16-pixel group that contains pixel v (arithmetic shift = floor for negatives).
__device__ __forceinline__ int tile_elem0(int v) { return (v >> 4) << 2; }

struct TileCoord {
  int f, s, ct;
  Roi roi;
  bool valid;
};

// tile id -> (frame, strip, column tile).  The ids a CTA visits advance by gridDim.x, so the frame index and the
// remainder are carried incrementally (TileCursor) instead of dividing per tile: the two integer divisions were 18 % of
// all instructions of this kernel in the first profile.
struct TileCursor {
  int f, rem;              // tile = f * per_frame + rem
  int step_f, step_rem;    // gridDim.x = step_f * per_frame + step_rem
  int per_frame;
  __device__ __forceinline__ void init(const K1Geom& g, int t, int stride) {
    per_frame = g.n_strips * g.n_ct;
    f = t / per_frame; rem = t - f * per_frame;
    step_f = stride / per_frame; step_rem = stride - step_f * per_frame;
  }
  __device__ __forceinline__ void advance() {
    f += step_f; rem += step_rem;
    if (rem >= per_frame) { rem -= per_frame; ++f; }
  }
};

__device__ __forceinline__ TileCoord decode_tile(const K1Geom& g, const TileCursor& cur) {
  TileCoord c;
  c.f = cur.f;
  if (g.n_ct == 1) { c.s = cur.rem; c.ct = 0; }
  else { c.s = cur.rem / g.n_ct; c.ct = cur.rem - c.s * g.n_ct; }
  c.roi = g.rois ? g.rois[c.f] : g.roi;
  c.valid = (c.s * kTileRows < c.roi.h) && (c.ct * g.tw_px < c.roi.w);
  return c;
}

// listed tile (K1Geom::tile_list): every listed tile is valid
__device__ __forceinline__ TileCoord decode_listed(const K1Geom& g, uint32_t e, bool want_roi) {
  TileCoord c;
  c.f = (int)(e >> 12); c.s = (int)((e >> 5) & 127u); c.ct = (int)(e & 31u);
  if (want_roi) c.roi = g.rois[c.f];
  c.valid = true;
  return c;
}

// One thread per frame appends the tiles its ROI touches to the work list (order irrelevant: every tile is independent).
__global__ void build_tile_list_kernel(const K1Geom g, uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
  pdl_enter();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= g.n_frames) return;
  const Roi r = g.rois[f];
  if (r.w <= 0 || r.h <= 0) return;
  const int ns = min((r.h + kTileRows - 1) / kTileRows, g.n_strips);
  const int nc = min((r.w + g.tw_px - 1) / g.tw_px, g.n_ct);
  uint32_t o = atomicAdd(count, (uint32_t)(ns * nc));
  for (int s = 0; s < ns; ++s)
    for (int ct = 0; ct < nc; ++ct) list[o++] = ((uint32_t)f << 12) | ((uint32_t)s << 5) | (uint32_t)ct;
}

cudaError_t launch_build_tile_list(const K1Geom& g, uint32_t* list, uint32_t* count, cudaStream_t st) {
  return launch_k(build_tile_list_kernel, (g.n_frames + 127) / 128, 128, 0, st, g, list, count);
}

// ------------------------------------------------------------------------------------------------
// K1a — scan_kernel: the streaming pass.  Persistent CTAs pull (frame, strip, column-tile) boxes through a TMA ring and
// only answer "which 32-bit words hold a byte above the threshold".  Tiles without such a word (the vast majority of an
// LED image) produce no output at all; for the others the CTA appends the word list to a global pool and files a HotTile
// record.  All exact arithmetic happens in K1c on those few tiles, with far more CTAs per SM than the 100 KB ring allows
// here — in the first (fused) version the rare slow path stalled the ring: 42 % of the warp stalls were barrier waits and
// the kernel sat at 0.49 of the HBM roofline.
// ------------------------------------------------------------------------------------------------
// Shared-memory layout (dynamic):  [0,64) mbarriers | misc[4] u32 | hot[kHotListCap] u16 | ring (1024-byte aligned)
constexpr int kHotListCap = 1024;        // hot words listed per tile; beyond that the tile is handed over as "dense"

__host__ __device__ inline size_t k1_ring_offset() {
  size_t off = 64 + 16 + (size_t)kHotListCap * 2;
  return (off + 1023) & ~(size_t)1023;
}

template <bool kLowThr, int kStages>
__global__ void __launch_bounds__(kK1Threads, 2) scan_kernel(const __grid_constant__ CUtensorMap tmap, const K1aArgs a) {
  pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const K1Geom& g = a.g;
  const int R = a.radius;
  const int kRows = kTileRows + 2 * R;
  const int box_w = g.box_w;                       // u32 per smem row
  const uint32_t stage_bytes = (uint32_t)(kRows * box_w * 4);          // bytes one TMA box delivers
  const uint32_t stage_stride = (stage_bytes + 127u) & ~127u;           // TMA destinations must be 128-byte aligned
  const int units_per_row = box_w >> 2;
  const int n_units = kRows * units_per_row;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem_raw + 64);         // [0] hot words of the current tile, [1] pool offset
  uint16_t* hot = reinterpret_cast<uint16_t*>(smem_raw + 80);
  uint8_t* ring = smem_raw + k1_ring_offset();

  const int tid = threadIdx.x;
  const bool listed = g.tile_list != nullptr;
  const int per_frame = g.n_strips * g.n_ct;
  const int n_tiles = listed ? (int)*g.tile_count : g.n_frames * per_frame;
  if (n_tiles == 0) return;                        // e.g. the whole-image retry pass of a tracking step that no stream needs
  if (tid == 0) {
    misc[0] = 0u; misc[1] = 0u;
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // producer state (thread 0 only)
  int ptile = blockIdx.x, pstage = 0;
  TileCursor pcur;
  pcur.init(g, blockIdx.x, gridDim.x);
  auto produce_one = [&]() {
    while (ptile < n_tiles) {
      TileCoord c = listed ? decode_listed(g, g.tile_list[ptile], true) : decode_tile(g, pcur);
      ptile += gridDim.x;
      pcur.advance();
      if (!c.valid) continue;
      int stage = pstage;
      pstage = (pstage + 1 == kStages) ? 0 : pstage + 1;
      int x_elem0 = tile_elem0(c.roi.x + c.ct * g.tw_px - R);
      int y0 = c.roi.y + c.s * kTileRows - R;
      mbar_expect_tx(&bars[stage], stage_bytes);
      tma_load_3d(ring + (size_t)stage * stage_stride, &tmap, &bars[stage], x_elem0, y0, g.frame_map ? g.frame_map[c.f] : c.f);
      return;
    }
  };
  if (tid == 0)
    for (int s = 0; s < kStages; ++s) produce_one();

  const uint32_t thr_k = (uint32_t)a.thr_k;
  auto push_unit = [&](int u, const uint4& v) {     // exact per-word test of one 16-byte unit
    const int row = u / units_per_row;
    const int w0 = (u - row * units_per_row) * 4;
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (any_byte_gt<kLowThr>(wv[q], thr_k)) {
        uint32_t slot = atomicAdd(&misc[0], 1u);
        if (slot < (uint32_t)kHotListCap) hot[slot] = (uint16_t)((row << 8) | (w0 + q));
      }
  };

  int cstage = 0;
  uint32_t cparity = 0;
  TileCursor ccur;
  ccur.init(g, blockIdx.x, gridDim.x);
  for (int idx = blockIdx.x; idx < n_tiles; idx += gridDim.x, ccur.advance()) {
    int tile = idx;                                  // id filed with a hot tile: f * per_frame + strip * n_ct + column tile
    if (listed) {
      const TileCoord c = decode_listed(g, g.tile_list[idx], false);
      tile = c.f * per_frame + c.s * g.n_ct + c.ct;
    } else {
      const TileCoord c = decode_tile(g, ccur);
      if (!c.valid) continue;
    }
    const int stage = cstage;
    const uint32_t parity = cparity;
    if (++cstage == kStages) { cstage = 0; cparity ^= 1u; }
    mbar_wait(&bars[stage], parity);
    const uint4* units = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_stride);

    // four 16-byte units (64 bytes) per test: OR the sixteen words, one SWAR compare; split up only when it is positive
    int pushed = 0;
    for (int base = 0; base < n_units; base += 4 * kK1Threads) {
      const int u0 = base + tid;
      uint4 v0 = (u0 < n_units) ? units[u0] : make_uint4(0, 0, 0, 0);
      uint4 v1 = (u0 + kK1Threads < n_units) ? units[u0 + kK1Threads] : make_uint4(0, 0, 0, 0);
      uint4 v2 = (u0 + 2 * kK1Threads < n_units) ? units[u0 + 2 * kK1Threads] : make_uint4(0, 0, 0, 0);
      uint4 v3 = (u0 + 3 * kK1Threads < n_units) ? units[u0 + 3 * kK1Threads] : make_uint4(0, 0, 0, 0);
      uint32_t o0 = (v0.x | v0.y) | (v0.z | v0.w), o1 = (v1.x | v1.y) | (v1.z | v1.w);
      uint32_t o2 = (v2.x | v2.y) | (v2.z | v2.w), o3 = (v3.x | v3.y) | (v3.z | v3.w);
      if (any_byte_gt<kLowThr>((o0 | o1) | (o2 | o3), thr_k)) {
        pushed = 1;
        if (any_byte_gt<kLowThr>(o0, thr_k)) push_unit(u0, v0);
        if (any_byte_gt<kLowThr>(o1, thr_k)) push_unit(u0 + kK1Threads, v1);
        if (any_byte_gt<kLowThr>(o2, thr_k)) push_unit(u0 + 2 * kK1Threads, v2);
        if (any_byte_gt<kLowThr>(o3, thr_k)) push_unit(u0 + 3 * kK1Threads, v3);
      }
    }
    // one barrier per tile: all reads of this stage are done, and the hot/cold decision is block-uniform (a thread that
    // races ahead into the next tile may already be pushing, so the shared counter itself must not be used to decide)
    if (__syncthreads_or(pushed)) {
      if (tid == 0) {
        const uint32_t n_hot = misc[0];
        uint32_t off = 0xffffffffu;                  // dense hand-over unless the list fits
        if (n_hot != 0u && n_hot <= (uint32_t)kHotListCap) {
          uint32_t o = atomicAdd(&a.counters[1], n_hot);
          if (o + n_hot <= a.pool_capacity) off = o;
        }
        misc[1] = off;
        misc[2] = n_hot;
        if (n_hot != 0u) {
          uint32_t rec = atomicAdd(&a.counters[0], 1u);
          a.hot_tiles[rec] = make_uint4((uint32_t)tile, (off == 0xffffffffu) ? 0xffffffffu : n_hot, off, 0u);
        }
      }
      __syncthreads();
      const uint32_t off = misc[1], n_hot = misc[2];
      if (off != 0xffffffffu)
        for (uint32_t i = tid; i < n_hot; i += kK1Threads) a.pool[off + i] = hot[i];
      __syncthreads();
      if (tid == 0) misc[0] = 0u;
      __syncthreads();
    }
    if (tid == 0) produce_one();
  }
}

// ------------------------------------------------------------------------------------------------
// K1c — blur_kernel: exact threshold + fixed-point Gaussian on the hot tiles only, reading the few source pixels it
// needs straight from global memory (about 1 % of the frame bytes), and writing the 1-bit foreground rows + row flags.
// ------------------------------------------------------------------------------------------------
// a hot tile offers only ~50-100 work items and six block barriers: small CTAs, many per SM (measured @8192 frames: 128 threads x 8
// CTAs/SM 0.203 ms, 64 x 16 0.179 ms, 32 x 24 0.257 ms)
#ifndef MPE_BLUR_THREADS
#define MPE_BLUR_THREADS 64
#endif
#ifndef MPE_BLUR_CTAS_PER_SM
#define MPE_BLUR_CTAS_PER_SM 16
#endif
constexpr int kBlurThreads = MPE_BLUR_THREADS;
constexpr int kBlurWarps = kBlurThreads / 32;

__device__ __forceinline__ int blur_smem_words(int box_w, int tw_px) {
  return 4 + kTileRows * ((box_w + 31) >> 5) + kTileRows * ((tw_px + 31) >> 5);
}

template <int RT>   // RT > 0: radius known at compile time (fully unrolled); RT == 0: a.radius, generic loops (R up to kMaxRadius)
__global__ void __launch_bounds__(kBlurThreads) blur_kernel(const K1aArgs a) {
  pdl_enter();
  const int R = (RT > 0) ? RT : a.radius;
  constexpr int kR = (RT > 0) ? RT : kMaxRadius;     // array bounds
  const int dj_max = (R + 3) >> 2;                    // source words that can influence an output word: j-dj_max .. j+dj_max
  extern __shared__ __align__(16) uint8_t bsm[];
  const K1Geom& g = a.g;
  const int box_w = g.box_w;
  const int hot_wpr = (box_w + 31) >> 5;
  const int om_wpr = (g.tw_px + 31) >> 5;
  uint32_t* misc = reinterpret_cast<uint32_t*>(bsm);                  // [1] #active words, [2] row mask
  uint32_t* act = misc + 4;                                            // [kTileRows][hot_wpr]   kept all-zero between tiles
  uint32_t* omask = act + kTileRows * hot_wpr;                         // [kTileRows][om_wpr]    kept all-zero between tiles
  uint16_t* list = reinterpret_cast<uint16_t*>(omask + kTileRows * om_wpr);   // [kTileRows * box_w]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < blur_smem_words(box_w, g.tw_px); i += kBlurThreads) misc[i] = 0u;
  __syncthreads();
  const uint32_t n_hot_tiles = a.counters[0];
  const uint32_t thr = (uint32_t)a.threshold;
  const int per_frame = g.n_strips * g.n_ct;

  for (uint32_t ht = blockIdx.x; ht < n_hot_tiles; ht += gridDim.x) {
    const uint4 rec = a.hot_tiles[ht];
    const int tile = (int)rec.x;
    TileCoord c;
    c.f = tile / per_frame;
    { int rem = tile - c.f * per_frame; c.s = rem / g.n_ct; c.ct = rem - c.s * g.n_ct; }
    c.roi = g.rois ? g.rois[c.f] : g.roi;
    const int x_elem0 = tile_elem0(c.roi.x + c.ct * g.tw_px - R);
    const int y0 = c.roi.y + c.s * kTileRows - R;
    const int out_rows = min(kTileRows, c.roi.h - c.s * kTileRows);
    const int tile_x0 = c.roi.x + c.ct * g.tw_px;                      // output pixel range of this tile
    const int tile_x1 = min(c.roi.x + c.roi.w, tile_x0 + g.tw_px);
    const uint8_t* frame = a.frames + (size_t)(g.frame_map ? g.frame_map[c.f] : c.f) * a.frame_stride;

    // (B) every hot source word marks the output words it can influence: rows row-2R..row, words j-1..j+1
    if (rec.y == 0xffffffffu) {                      // dense hand-over: every word of the tile is a candidate
      for (int item = tid; item < out_rows * hot_wpr; item += kBlurThreads) {
        int hw = item % hot_wpr;
        int nb = min(32, box_w - hw * 32);
        act[item] = (nb >= 32) ? 0xffffffffu : ((1u << nb) - 1u);
      }
    } else {
      for (uint32_t t = tid; t < rec.y; t += kBlurThreads) {
        const uint32_t e = a.pool[rec.z + t];
        const int row = (int)(e >> 8), j = (int)(e & 0xff);
        const int r_lo = max(row - 2 * R, 0), r_hi = min(row, out_rows - 1);
        for (int jj = max(j - dj_max, 0); jj <= min(j + dj_max, box_w - 1); ++jj) {
          const uint32_t bit = 1u << (jj & 31);
          for (int r = r_lo; r <= r_hi; ++r) atomicOr(&act[r * hot_wpr + (jj >> 5)], bit);
        }
      }
    }
    __syncthreads();
    // (C) compact the marked words; the bitmap is cleared on the way
    for (int item = tid; item < out_rows * hot_wpr; item += kBlurThreads) {
      uint32_t d = act[item];
      if (!d) continue;
      act[item] = 0u;
      const int r = item / hot_wpr, hw = item - r * hot_wpr;
      while (d) {
        const int bit = __ffs(d) - 1;
        d &= d - 1;
        list[atomicAdd(&misc[1], 1u)] = (uint16_t)((r << 8) | (hw * 32 + bit));
      }
    }
    __syncthreads();
    const int n_list = (int)misc[1];
    // (D) exact arithmetic: cv::threshold(THRESH_TOZERO) -> 8.8 horizontal pass -> 16.16 vertical pass -> round, != 0 ?
    for (int e = tid; e < n_list; e += kBlurThreads) {
      const int r = list[e] >> 8, j = list[e] & 0xff;
      const int first_px = 4 * x_elem0 + 4 * j;                        // image x of the word's first pixel
      if (first_px + 3 < tile_x0 || first_px >= tile_x1) continue;
      const int Y = y0 + R + r;                                        // image row of this output row
      const bool interior = (first_px - R >= c.roi.x) && (first_px + 3 + R < c.roi.x + c.roi.w) &&
                            (Y - R >= c.roi.y) && (Y + R < c.roi.y + c.roi.h);
      uint32_t hsum[2 * kR + 1][4];
#pragma unroll
      for (int dr = 0; dr <= 2 * R; ++dr) {
        uint32_t px[4 + 2 * kR];
        if (interior) {
          const uint8_t* src = frame + (size_t)(Y + dr - R) * a.pitch + (first_px - R);
#pragma unroll
          for (int k = 0; k < 4 + 2 * R; ++k) px[k] = __ldg(src + k);
        } else {
          const int yy = c.roi.y + reflect101(Y + dr - R - c.roi.y, c.roi.h);
          const uint8_t* srow = frame + (size_t)yy * a.pitch;
#pragma unroll
          for (int k = 0; k < 4 + 2 * R; ++k) {
            const int xx = c.roi.x + reflect101(first_px + k - R - c.roi.x, c.roi.w);
            px[k] = __ldg(srow + xx);
          }
        }
#pragma unroll
        for (int k = 0; k < 4 + 2 * R; ++k) px[k] = (px[k] > thr) ? px[k] : 0u;      // strict >
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t acc = 0;
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc += a.taps[k] * px[q + k];
          hsum[dr][q] = acc;
        }
      }
      uint32_t bits = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int X = first_px + q;
        uint32_t acc = 0;
#pragma unroll
        for (int dr = 0; dr <= 2 * R; ++dr) acc += a.taps[dr] * hsum[dr][q];
        if (X >= tile_x0 && X < tile_x1 && ((acc + 32768u) >> 16)) bits |= 1u << q;
      }
      if (bits) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (bits & (1u << q)) {
            const int bx = first_px + q - tile_x0;
            atomicOr(&omask[r * om_wpr + (bx >> 5)], 1u << (bx & 31));
          }
        atomicOr(&misc[2], 1u << r);
      }
    }
    __syncthreads();
    // (E) rows with foreground go to the global mask; rows without are not written (their flag bit stays 0)
    const uint32_t rowmask = misc[2];
    __syncthreads();
    if (tid == 0) {
      a.rowflags[(size_t)c.f * g.flags_per_frame + c.s * g.n_ct + c.ct] = rowmask;
      misc[1] = 0u; misc[2] = 0u;
    }
    const int words_per_ct = g.tw_px >> 5;   // only used when n_ct > 1 (tw_px is a multiple of 32 then)
    // warp w copies the rows r with r % kBlurWarps == w
    constexpr uint32_t kRowsOfWarp0 = (kBlurWarps == 1) ? 0xffffffffu : (kBlurWarps == 2) ? 0x55555555u : (kBlurWarps == 4) ? 0x11111111u : 0x01010101u;
    uint32_t rm = rowmask & (kRowsOfWarp0 << warp);
    while (rm) {
      const int r = __ffs(rm) - 1;
      rm &= rm - 1;
      uint32_t* dst = a.mask + ((size_t)c.f * g.mask_rows + (c.s * kTileRows + r)) * g.mask_wpr + c.ct * words_per_ct;
      for (int w = lane; w < om_wpr; w += 32) {
        dst[w] = omask[r * om_wpr + w];
        omask[r * om_wpr + w] = 0u;
      }
    }
    __syncthreads();
  }
}

size_t find_leds_smem_bytes(const K1Geom& g, int radius, int stages) {
  int rows = kTileRows + 2 * radius;
  size_t stage_stride = ((size_t)rows * g.box_w * 4 + 127) & ~(size_t)127;
  return k1_ring_offset() + (size_t)stages * stage_stride;
}

template <bool kLow, int kStages>
static cudaError_t launch_scan_inst(const K1aArgs& a, const CUtensorMap& tmap, int n_sms, cudaStream_t st) {
  size_t smem = find_leds_smem_bytes(a.g, a.radius, kStages);
  auto kern = scan_kernel<kLow, kStages>;
  static SmemAttrCache configured;
  {
    cudaError_t e = ensure_dynamic_smem(kern, smem, configured, 0);
    if (e != cudaSuccess) return e;
  }
  int n_tiles = a.g.n_frames * a.g.n_strips * a.g.n_ct;
  // CTAs per SM: two for the ~100 KB rings of whole-image rows; up to four (the register limit: 256 threads x 60 registers) for the
  // small boxes of tracking-mode ROIs, whose loads are latency bound — more boxes in flight per SM
  int ctas_per_sm = (int)((220 * 1024) / smem);
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int grid = n_tiles < n_sms * ctas_per_sm ? n_tiles : n_sms * ctas_per_sm;
  if (grid < 1) grid = 1;
  return launch_k(kern, grid, kK1Threads, smem, st, tmap, a);
}

static cudaError_t launch_scan(const K1aArgs& a, const CUtensorMap& tmap, int n_sms, cudaStream_t st) {
  // Ring depth.  Per tile a CTA runs a latency chain (mbarrier wait -> shared-memory reads -> block barrier), so what hides latency
  // is CTAs per SM, not a deeper ring: measured @8192 x 752x480: 2 CTAs x 3 stages 0.521 ms (0.87 of the HBM peak), 3 CTAs x 2
  // stages 0.466 ms (0.97), 1 CTA x 4 stages 0.939 ms.  Rule: the depth (>= 2) that lets the most CTAs share an SM (at most four:
  // 256 threads x 60 registers), ties to the deeper ring.  MPE_SCAN_STAGES=2|3|4 forces a depth.
  const int R = a.radius;
  bool low = a.threshold < 128;
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MPE_SCAN_STAGES"); forced = e ? atoi(e) : 0; }
  int stages = 2, best_ctas = 0;
  for (int sdepth = 2; sdepth <= 4; ++sdepth) {
    int ctas = (int)((220 * 1024) / find_leds_smem_bytes(a.g, R, sdepth));
    if (ctas > 4) ctas = 4;
    if (ctas >= best_ctas && ctas >= 1) { best_ctas = ctas; stages = sdepth; }
  }
  if (forced >= 1 && forced <= 4) stages = forced;
  if (stages == 1) return low ? launch_scan_inst<true, 1>(a, tmap, n_sms, st) : launch_scan_inst<false, 1>(a, tmap, n_sms, st);
  if (stages == 4) return low ? launch_scan_inst<true, 4>(a, tmap, n_sms, st) : launch_scan_inst<false, 4>(a, tmap, n_sms, st);
  if (stages == 3) return low ? launch_scan_inst<true, 3>(a, tmap, n_sms, st) : launch_scan_inst<false, 3>(a, tmap, n_sms, st);
  return low ? launch_scan_inst<true, 2>(a, tmap, n_sms, st) : launch_scan_inst<false, 2>(a, tmap, n_sms, st);
}

template <int RT>
static cudaError_t launch_blur_r(const K1aArgs& a, int n_sms, cudaStream_t st) {
  size_t bsmem = (size_t)(4 + kTileRows * ((a.g.box_w + 31) >> 5) + kTileRows * ((a.g.tw_px + 31) >> 5)) * 4 + (size_t)kTileRows * a.g.box_w * 2;
  int n_tiles = a.g.n_frames * a.g.n_strips * a.g.n_ct;
  int per_sm = MPE_BLUR_CTAS_PER_SM;
  if ((size_t)per_sm * (bsmem + 1024) > 220 * 1024) per_sm = (int)((220 * 1024) / (bsmem + 1024));
  if (per_sm < 1) per_sm = 1;
  int grid = n_tiles < n_sms * per_sm ? n_tiles : n_sms * per_sm;
  return launch_k(blur_kernel<RT>, grid, kBlurThreads, bsmem, st, a);
}

cudaError_t launch_blur_tiles(const K1aArgs& a, int radius, int n_sms, cudaStream_t st) {
  switch (radius) {
    case 1: return launch_blur_r<1>(a, n_sms, st);
    case 2: return launch_blur_r<2>(a, n_sms, st);
    case 3: return launch_blur_r<3>(a, n_sms, st);
    case 4: return launch_blur_r<4>(a, n_sms, st);
    default: return (radius >= 1 && radius <= kMaxRadius) ? launch_blur_r<0>(a, n_sms, st) : cudaErrorInvalidValue;
  }
}

cudaError_t launch_find_leds(const K1aArgs& a, const CUtensorMap& tmap, int radius, int n_sms, cudaStream_t st) {
  if (radius < 1 || radius > kMaxRadius || radius != a.radius) return cudaErrorInvalidValue;
  return launch_scan(a, tmap, n_sms, st);
}

// ------------------------------------------------------------------------------------------------
// K1b — contours, moments, filters, undistortion.  One warp per frame.
// ------------------------------------------------------------------------------------------------
struct MaskView {
  const uint32_t* flags;   // this frame's row-flag words
  const uint32_t* mask;    // this frame's mask rows
  int w, h;                // ROI size
  int n_ct, roi_n_ct, tw_px, wpr, words_per_ct;
  const uint32_t* win;     // optional shared-memory copy of the ROI's mask rows (h x win_wpr words, rows without foreground zeroed)
  int win_wpr;
};

// kWin: the launch holds the ROI's mask rows in shared memory (m.win valid for every frame-warp that takes this path)
template <bool kWin>
__device__ __forceinline__ uint32_t mv_word(const MaskView& m, int y, int wi) {
  // word wi of ROI row y, 0 where the producing tile reported no foreground in that row
  if (kWin) {
    if ((unsigned)y >= (unsigned)m.h || (unsigned)wi >= (unsigned)m.win_wpr) return 0u;
    return m.win[y * m.win_wpr + wi];
  }
  if ((unsigned)y >= (unsigned)m.h || wi < 0 || wi >= m.wpr) return 0u;
  int ct = (m.roi_n_ct == 1) ? 0 : min(wi / m.words_per_ct, m.roi_n_ct - 1);
  uint32_t fl = m.flags[(y >> 5) * m.n_ct + ct];
  if (!((fl >> (y & 31)) & 1u)) return 0u;
  return __ldg(m.mask + (size_t)y * m.wpr + wi);
}
template <bool kWin>
__device__ __forceinline__ bool mv_bit(const MaskView& m, int x, int y) {
  if ((unsigned)x >= (unsigned)m.w || (unsigned)y >= (unsigned)m.h) return false;
  if (kWin) return (m.win[y * m.win_wpr + (x >> 5)] >> (x & 31)) & 1u;      // x < w implies x >> 5 < win_wpr
  return (mv_word<false>(m, y, x >> 5) >> (x & 31)) & 1u;
}

// direction codes as OpenCV: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE
// (two bits per direction, value + 1, packed: dx = 1,1,0,-1,-1,-1,0,1 and dy = 0,-1,-1,-1,0,1,1,1)
__device__ __forceinline__ int dir_dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }

struct Contour {
  long long a00, a10, a01;
  int minx, maxx, miny, maxy;
  int status;   // 1 accepted (true outer border of a component), 0 rejected, -1 step limit hit
};

// Follow the border that starts at (x0,y0) (a pixel whose W, NW, N, NE neighbours are background) exactly as
// icvFetchContour does for an outer border.  The candidate is rejected as soon as the border reaches a pixel
// that precedes the start in raster order: then (x0,y0) is not the first pixel of its component (the border
// is either the component's outer border seen from a later "tip", or the border of a hole).
// When (px,py) is given (px >= 0) the even-odd crossing number of the closed polygon with respect to that
// lattice point is returned in *inside.
template <bool kWin>
__device__ Contour trace_border(const MaskView& m, int x0, int y0, int px, int py, bool* inside) {
  Contour c;
  c.a00 = c.a10 = c.a01 = 0;
  c.minx = c.maxx = x0;
  c.miny = c.maxy = y0;
  c.status = 1;
  bool in = false;
  int s = 4, s_end = 4;
  do {
    s = (s - 1) & 7;
    if (mv_bit<kWin>(m, x0 + dir_dx(s), y0 + dir_dy(s))) break;
  } while (s != s_end);
  if (s == s_end) {   // isolated pixel: one-point contour, all sums zero
    if (inside) *inside = false;
    return c;
  }
  const int i1x = x0 + dir_dx(s), i1y = y0 + dir_dy(s);
  int cx = x0, cy = y0;           // i3
  int prevx = 0, prevy = 0;       // previously emitted point
  bool have_prev = false;
  const long long step_limit = 4ll * m.w * m.h + 16;
  long long steps = 0;
  for (;;) {
    int nx, ny;
    do {
      s = (s + 1) & 7;
      nx = cx + dir_dx(s);
      ny = cy + dir_dy(s);
    } while (!mv_bit<kWin>(m, nx, ny));
    // emit (cx,cy)
    if (have_prev) {
      long long dxy = (long long)prevx * cy - (long long)cx * prevy;
      c.a00 += dxy;
      c.a10 += dxy * (prevx + cx);
      c.a01 += dxy * (prevy + cy);
      if (px >= 0 && ((prevy > py) != (cy > py))) {
        int xi = (prevy == py) ? prevx : cx;     // the endpoint lying on row py
        if (px < xi) in = !in;
      }
    }
    prevx = cx; prevy = cy; have_prev = true;
    c.minx = min(c.minx, cx); c.maxx = max(c.maxx, cx);
    c.miny = min(c.miny, cy); c.maxy = max(c.maxy, cy);
    if (nx == x0 && ny == y0 && cx == i1x && cy == i1y) break;
    if (ny < y0 || (ny == y0 && nx < x0)) { c.status = 0; break; }
    if (++steps > step_limit) { c.status = -1; break; }
    cx = nx; cy = ny;
    s = (s + 4) & 7;
  }
  if (c.status == 1) {   // closing edge: last emitted point -> start
    long long dxy = (long long)prevx * y0 - (long long)x0 * prevy;
    c.a00 += dxy;
    c.a10 += dxy * (prevx + x0);
    c.a01 += dxy * (prevy + y0);
    if (px >= 0 && ((prevy > py) != (y0 > py))) {
      int xi = (prevy == py) ? prevx : x0;
      if (px < xi) in = !in;
    }
  }
  if (inside) *inside = in;
  return c;
}

// candidate starts in word wi of row y: foreground pixels whose W, NW, N and NE neighbours are background
template <bool kWin>
__device__ __forceinline__ uint32_t candidate_bits(const MaskView& m, int y, int wi) {
  uint32_t cur = mv_word<kWin>(m, y, wi);
  if (!cur) return 0u;
  uint32_t curL = mv_word<kWin>(m, y, wi - 1);
  uint32_t up = mv_word<kWin>(m, y - 1, wi), upL = mv_word<kWin>(m, y - 1, wi - 1), upR = mv_word<kWin>(m, y - 1, wi + 1);
  uint32_t Wn = (cur << 1) | (curL >> 31);
  uint32_t NW = (up << 1) | (upL >> 31);
  uint32_t NE = (up >> 1) | (upR << 31);
  return cur & ~Wn & ~up & ~NW & ~NE;
}

// Is the component starting at (px,py) enclosed by another component?  (RETR_EXTERNAL drops it then.)
// Fast exits: a clear axis ray from the start pixel to the ROI edge proves it is not enclosed.
template <bool kWin>
__device__ bool is_enclosed(const MaskView& m, int px, int py) {
  // left / right rays in the same row
  {
    bool blocked_l = false, blocked_r = false;
    int wi0 = px >> 5, b = px & 31;
    for (int wi = 0; wi <= wi0 && !blocked_l; ++wi) {
      uint32_t v = mv_word<kWin>(m, py, wi);
      if (wi == wi0) v &= (b == 0) ? 0u : (0xffffffffu >> (32 - b));
      blocked_l = v != 0;
    }
    if (!blocked_l) return false;
    // right ray, word-wise: skip the component's own run that starts at px, then look for any further foreground
    bool in_run = true;
    for (int wi = wi0; wi < m.wpr && !blocked_r; ++wi) {
      uint32_t v = mv_word<kWin>(m, py, wi);
      int lo = (wi == wi0) ? b : 0;                       // first bit of interest in this word
      v = (lo ? (v >> lo) : v);
      int nbits = 32 - lo;
      if (in_run) {
        uint32_t inv = ~v;
        if (nbits < 32) inv &= (1u << nbits) - 1u;
        if (inv == 0u) continue;                          // run covers the rest of this word
        int run = __ffs(inv) - 1;
        v = (run < 32) ? (v >> run) : 0u;
        in_run = false;
      }
      blocked_r = v != 0u;
    }
    if (!blocked_r) return false;
  }
  {
    // upwards: only rows whose row flag says "contains foreground" can block (mv_word returns 0 for the others)
    bool blocked_u = false;
    if (py > 0) {
      const int wi = px >> 5;
      const int ct = (m.roi_n_ct == 1) ? 0 : min(wi / m.words_per_ct, m.roi_n_ct - 1);
      for (int s = (py - 1) >> 5; s >= 0 && !blocked_u; --s) {
        uint32_t fl = m.flags[s * m.n_ct + ct];
        if (s == ((py - 1) >> 5)) fl &= 0xffffffffu >> (31 - ((py - 1) & 31));     // rows of this strip above py
        while (fl && !blocked_u) {
          const int r = 31 - __clz(fl);
          fl &= ~(1u << r);
          blocked_u = mv_bit<kWin>(m, px, s * kTileRows + r);
        }
      }
    }
    if (!blocked_u) return false;
    // downwards the ray first leaves through the component's own pixels; a later foreground pixel may still
    // belong to the component itself, so a blocked ray proves nothing — fall through to the exact test.
  }
  // exact test: any component whose first pixel precedes row py and whose outer polygon contains (px,py)
  const int strips = (m.h + kTileRows - 1) / kTileRows;
  for (int s = 0; s < strips && s * kTileRows < py; ++s) {
    uint32_t fl = 0;
    for (int ct = 0; ct < m.roi_n_ct; ++ct) fl |= m.flags[s * m.n_ct + ct];
    while (fl) {
      int r = __ffs(fl) - 1;
      fl &= fl - 1;
      int y = s * kTileRows + r;
      if (y >= py) break;
      for (int wi = 0; wi < m.wpr; ++wi) {
        uint32_t cand = candidate_bits<kWin>(m, y, wi);
        while (cand) {
          int b = __ffs(cand) - 1;
          cand &= cand - 1;
          bool inside = false;
          Contour c = trace_border<kWin>(m, wi * 32 + b, y, px, py, &inside);
          if (c.status == 1 && inside) return true;
        }
      }
    }
  }
  return false;
}

// cv::undistortPoints(src, dst, K, D, noArray(), P = K), default criteria (5 fixed-point iterations)
__device__ __forceinline__ void undistort_point(const DevCamera& cam, float sx, float sy, float* ox, float* oy) {
  const double fx = cam.K[0], fy = cam.K[4], cx = cam.K[2], cy = cam.K[5];
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = sx, y = sy;
  const double u = x, v = y;
  x = (x - cx) * ifx;
  y = (y - cy) * ify;
  if (cam.nD > 0) {
    const double* k = cam.D;
    double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      double r2 = x * x + y * y;
      double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) {
        x = (u - cx) * ifx;
        y = (v - cy) * ify;
        break;
      }
      double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
  }
  double xx = cam.K[0] * x + cam.K[1] * y + cam.K[2];
  double yy = cam.K[3] * x + cam.K[4] * y + cam.K[5];
  double ww = 1. / (cam.K[6] * x + cam.K[7] * y + cam.K[8]);
  *ox = (float)(xx * ww);
  *oy = (float)(yy * ww);
}

constexpr int kBlobWarpsPerCta = 4;
// kMaxFlagWords (mpe_internal.cuh): row-flag words cached per frame, >= n_strips * n_ct of the launch (the host rejects larger
// geometries).  1080p with the 256-px tracking tiles needs 34 x 8 = 272; the first limit of 160 was found by the 1080p tracking test:
// out-of-range flag words fed the border follower garbage and its step limit of 4*w*h turned that into a hang.

// Per-warp scratch.  The two variable parts are sized per launch (row flags: strips x column tiles of the geometry; row list: image
// rows), so that small images keep the footprint small and more frame-warps share an SM.
struct KeptList {                  // the blobs of one frame that passed the filters
  int n_kept;
  int flags;
  int kept_key[MPE_MAX_BLOBS];     // raster index of the contour start (sort key)
  float kept_cx[MPE_MAX_BLOBS];
  float kept_cy[MPE_MAX_BLOBS];
};
struct WarpScratch {
  int cand_x[kCandCap];
  int cand_y[kCandCap];
  int n_cand;
  int n_rows;
  KeptList kl;
  uint32_t* rowflags;              // [flags_cap]
  uint16_t* rows;                  // [rows_cap] rows of this frame that contain foreground
  int flags_cap, rows_cap;
};

__host__ __device__ inline size_t k1b_scratch_stride(int flags_cap, int rows_cap, int win_words = 0) {
  size_t n = sizeof(WarpScratch) + (size_t)flags_cap * 4 + (((size_t)rows_cap * 2 + 3) & ~(size_t)3) + (size_t)win_words * 4;
  return (n + 15) & ~(size_t)15;
}
// Small launches (a handful of cameras: the latency case) copy the ROI's mask rows into shared memory first: the border follower
// is a chain of dependent single-word reads, ~10x cheaper from shared memory than from L2.  Large batches keep the occupancy instead.
constexpr int kK1bWindowMaxFrames = 64;
constexpr int kK1bPoolMinFrames = 2048;    // from here on a warp serves four frames (extract_blobs_pooled_kernel)
constexpr int kK1bWindowWords = 8192;       // 32 KB per frame-warp: e.g. a 256 x 1024 px ROI; larger ROIs read the global mask

// One contour-start candidate: follow its border, drop holes / enclosed components, apply the blob filters (led_detector.cpp:67-81)
template <bool kWin>
__device__ __forceinline__ void process_one_candidate(const MaskView& m, const K1bArgs& a, const Roi roi, KeptList& kl, int x0, int y0) {
  Contour c = trace_border<kWin>(m, x0, y0, -1, -1, nullptr);
  if (c.status == -1) atomicOr(&kl.flags, MPE_F_TRACE_ABORT);
  if (c.status == 1 && !is_enclosed<kWin>(m, x0, y0)) {
    double area = fabs((double)c.a00) * 0.5;                       // cv::contourArea
    int rw = c.maxx - c.minx + 1, rh = c.maxy - c.miny + 1;      // cv::boundingRect
    // cv::moments(contour): m00 = a00 * (+-0.5), m10 = a10 * (+-1/6), m01 = a01 * (+-1/6), sign so that m00 > 0
    double m00 = 0, m10 = 0, m01 = 0;
    if (c.a00 != 0) {
      double db1_2 = (c.a00 > 0) ? 0.5 : -0.5;
      double db1_6 = (c.a00 > 0) ? 0.16666666666666666666666666666667 : -0.16666666666666666666666666666667;
      m00 = (double)c.a00 * db1_2;
      m10 = (double)c.a10 * db1_6;
      m01 = (double)c.a01 * db1_6;
    }
    float mcx = (float)(m10 / m00) + (float)roi.x;                  // Point2f(m10/m00, m01/m00) + Point2f(ROI.x, ROI.y)
    float mcy = (float)(m01 / m00) + (float)roi.y;
    const double pi = 3.1415926535897932384626433832795;
    double wh = fabs(1 - fmin((double)rw / (double)rh, (double)rh / (double)rw));
    double hw2 = (double)(rw / 2), hh2 = (double)(rh / 2);          // integer division, led_detector.cpp:80-81
    double cw = fabs(1 - (area / (pi * (hw2 * hw2))));
    double ch = fabs(1 - (area / (pi * (hh2 * hh2))));
    bool keep = area >= a.bp.min_blob_area && area <= a.bp.max_blob_area && wh <= a.bp.max_width_height_distortion &&
                cw <= a.bp.max_circular_distortion && ch <= a.bp.max_circular_distortion;
    if (keep) {
      int slot = atomicAdd(&kl.n_kept, 1);
      if (slot < MPE_MAX_BLOBS) {
        kl.kept_key[slot] = y0 * m.w + x0;
        kl.kept_cx[slot] = mcx;
        kl.kept_cy[slot] = mcy;
      } else {
        atomicOr(&kl.flags, MPE_F_BLOB_OVERFLOW);
      }
    }
  }
}

template <bool kWin>
__device__ void process_candidates(const MaskView& m, const K1bArgs& a, const Roi roi, WarpScratch& ws, int lane) {
  __syncwarp();
  int n = min(ws.n_cand, kCandCap);
  for (int base = 0; base < n; base += 32) {
    int i = base + lane;
    if (i < n) process_one_candidate<kWin>(m, a, roi, ws.kl, ws.cand_x[i], ws.cand_y[i]);
  }
  __syncwarp();
  if (lane == 0) ws.n_cand = 0;
  __syncwarp();
}

// kept blob i of a frame: its place in cv::findContours' output order (reverse raster order of the contour starts) and its
// undistorted position (led_detector.cpp:97-110)
__device__ __forceinline__ void emit_blob(const K1bArgs& a, const KeptList& kl, int f, int i, int n) {
  const int key = kl.kept_key[i];
  int rank = 0;
  for (int j = 0; j < n; ++j) rank += (kl.kept_key[j] > key) ? 1 : 0;
  float ux, uy;
  undistort_point(a.cam, kl.kept_cx[i], kl.kept_cy[i], &ux, &uy);
  const size_t o = ((size_t)f * MPE_MAX_BLOBS + rank) * 2;
  a.centers[o] = kl.kept_cx[i];
  a.centers[o + 1] = kl.kept_cy[i];
  a.det[o] = (double)ux;
  a.det[o + 1] = (double)uy;
}

// The rows of a frame that contain foreground, from its row flags (order irrelevant: the blobs are sorted at the end)
__device__ __forceinline__ void list_rows(const uint32_t* rowflags, int roi_strips, int n_ct, int roi_n_ct, int* n_rows, uint16_t* rows, int rows_cap, int lane) {
  for (int s = lane; s < roi_strips; s += 32) {
    uint32_t fl = 0;
    for (int ct = 0; ct < roi_n_ct; ++ct) fl |= rowflags[s * n_ct + ct];
    while (fl) {
      int r = __ffs(fl) - 1;
      fl &= fl - 1;
      int slot = atomicAdd(n_rows, 1);
      if (slot < rows_cap) rows[slot] = (uint16_t)(s * kTileRows + r);
    }
  }
}

// contour-start candidates: one lane per foreground row, then lane-parallel border following
template <bool kWin>
__device__ __forceinline__ void find_candidates_and_trace(const MaskView& m, const K1bArgs& a, const Roi roi, WarpScratch& ws, int lane, int roi_wpr) {
  const int rows_cap = ws.rows_cap;
  const int n_rows = min(ws.n_rows, rows_cap);
  for (int base = 0; base < n_rows; base += 32) {
    if (base + lane < n_rows) {
      const int y = ws.rows[base + lane];
      // non-zero words of the row first (independent loads), then the neighbourhood test only where needed
      unsigned long long nz = 0;
      for (int w0 = 0; w0 < roi_wpr; w0 += 8) {                // eight loads in flight; a branch per loaded word would serialise them
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (w0 + u < roi_wpr) ? mv_word<kWin>(m, y, w0 + u) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) nz |= (unsigned long long)(v[u] != 0u) << ((w0 + u) & 63);
      }
      for (int wi = 0; wi < roi_wpr; ++wi) {
        if (roi_wpr <= 64 && !((nz >> wi) & 1ull)) continue;
        uint32_t cand = candidate_bits<kWin>(m, y, wi);
        while (cand) {
          int bbit = __ffs(cand) - 1;
          cand &= cand - 1;
          int slot = atomicAdd(&ws.n_cand, 1);
          if (slot < kCandCap) { ws.cand_x[slot] = wi * 32 + bbit; ws.cand_y[slot] = y; }
          else atomicOr(&ws.kl.flags, MPE_F_BLOB_OVERFLOW);
        }
      }
    }
    __syncwarp();
    if (ws.n_cand >= kCandCap / 2) process_candidates<kWin>(m, a, roi, ws, lane);
  }
  if (ws.n_cand > 0) process_candidates<kWin>(m, a, roi, ws, lane);
  __syncwarp();
}

// latency bound (dependent mask loads while following a border): occupancy pays.  Eight CTAs per SM = 64 registers per thread and
// a per-launch sized scratch; measured @8192 frames: 103 registers / 4 CTAs 0.271 ms, 80 / 6 0.220 ms, 64 / 8 0.195 ms
#ifndef MPE_K1B_MINBLOCKS
#define MPE_K1B_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(32 * kBlobWarpsPerCta, MPE_K1B_MINBLOCKS) extract_blobs_kernel(const K1bArgs a, int win_words) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t k1b_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kBlobWarpsPerCta + warp;
  if (f >= a.g.n_frames) return;
  if (a.active && !a.active[f]) return;
  const K1Geom& g = a.g;
  const int flags_cap = min(g.flags_per_frame, kMaxFlagWords), rows_cap = g.mask_rows;
  uint8_t* my = k1b_smem + (size_t)warp * k1b_scratch_stride(flags_cap, rows_cap, win_words);
  WarpScratch& ws = *reinterpret_cast<WarpScratch*>(my);
  if (lane == 0) {
    ws.rowflags = reinterpret_cast<uint32_t*>(my + sizeof(WarpScratch));
    ws.rows = reinterpret_cast<uint16_t*>(my + sizeof(WarpScratch) + (size_t)flags_cap * 4);
    ws.flags_cap = flags_cap; ws.rows_cap = rows_cap;
  }
  __syncwarp();
  const Roi roi = g.rois ? g.rois[f] : g.roi;
  const int roi_wpr = (roi.w + 31) >> 5;
  const int roi_n_ct = (roi.w + g.tw_px - 1) / g.tw_px;
  const int roi_strips = (roi.h + kTileRows - 1) / kTileRows;

  // ---- cache the row flags of this frame; list the rows that contain foreground (order irrelevant: sorted at the end)
  if (lane == 0) { ws.n_cand = 0; ws.kl.n_kept = 0; ws.kl.flags = 0; ws.n_rows = 0; }
  __syncwarp();
  const uint32_t* gflags = a.rowflags + (size_t)f * g.flags_per_frame;
  for (int i = lane; i < roi_strips * g.n_ct && i < flags_cap; i += 32) ws.rowflags[i] = gflags[i];
  __syncwarp();
  list_rows(ws.rowflags, roi_strips, g.n_ct, roi_n_ct, &ws.n_rows, ws.rows, rows_cap, lane);
  __syncwarp();

  MaskView m;
  m.flags = ws.rowflags;
  m.mask = a.mask + (size_t)f * g.mask_rows * g.mask_wpr;
  m.w = roi.w; m.h = roi.h;
  m.n_ct = g.n_ct;                 // row flags are stored with the launch-wide n_ct stride
  m.roi_n_ct = roi_n_ct;
  m.tw_px = g.tw_px;
  m.wpr = g.mask_wpr;
  m.words_per_ct = g.tw_px >> 5;
  m.win = nullptr; m.win_wpr = 0;
  if (win_words > 0 && roi.h * roi_wpr <= win_words && ws.n_rows <= rows_cap) {
    uint32_t* win = reinterpret_cast<uint32_t*>(my + sizeof(WarpScratch) + (size_t)flags_cap * 4 + (((size_t)rows_cap * 2 + 3) & ~(size_t)3));
    for (int i = lane; i < roi.h * roi_wpr; i += 32) win[i] = 0u;
    __syncwarp();
    const int total = ws.n_rows * roi_wpr;
    for (int base = 0; base < total; base += 32 * 8) {          // eight independent loads per lane in flight, then the stores
      uint32_t v[8];
      int at[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * 32 + lane;
        at[u] = -1; v[u] = 0u;
        if (i < total) {
          const int y = ws.rows[i / roi_wpr], wi = i - (i / roi_wpr) * roi_wpr;
          at[u] = y * roi_wpr + wi;
          v[u] = mv_word<false>(m, y, wi);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (at[u] >= 0) win[at[u]] = v[u];
    }
    __syncwarp();
    m.win = win; m.win_wpr = roi_wpr;
  }

  if (m.win) find_candidates_and_trace<true>(m, a, roi, ws, lane, roi_wpr);
  else find_candidates_and_trace<false>(m, a, roi, ws, lane, roi_wpr);

  // ---- order: cv::findContours returns the contours in reverse raster order of their start pixels ----
  const int n = min(ws.kl.n_kept, MPE_MAX_BLOBS);
  for (int i = lane; i < n; i += 32) emit_blob(a, ws.kl, f, i, n);
  if (lane == 0) {
    a.n_det[f] = n;
    a.flags[f] = ws.kl.flags;
  }
}

// ------------------------------------------------------------------------------------------------
// K1b for large batches: one warp per GROUP of kPoolFrames frames.
// A frame has ~5 contour-start candidates, so a warp that follows the borders of its own frame keeps ~6 of 32 lanes busy in
// that phase and in the undistortion (measured @8192 frames: 6.7 active lanes per instruction, 133 M warp instructions, the
// kernel latency bound with every resident warp working).  Pooling the candidates of a CTA's four frame-warps behind a block
// barrier halved the instructions but left three of four warps waiting (0.204 -> 0.189 ms only).  Here every warp stays busy:
// it lists rows and candidates frame by frame (lane = row), then follows the pooled candidates of its four frames with a lane
// each, then undistorts the pooled blobs.  Same functions per candidate as the kernel above, hence identical results.
// ------------------------------------------------------------------------------------------------
// (frames per warp <= 4: two tag bits in the candidate word)
template <int kPoolFrames>
struct PoolScratch {
  uint32_t cand[kCandCap];         // frame slot << 30 | y << 15 | x
  int n_cand;
  int n_rows;
  KeptList kl[kPoolFrames];
  MaskView mv[kPoolFrames];        // row flags read from global memory (the shared-memory copy is reused by the next frame)
  Roi roi[kPoolFrames];
};
template <int kPoolFrames>
__host__ __device__ inline size_t k1b_pool_stride(int flags_cap, int rows_cap) {
  size_t n = sizeof(PoolScratch<kPoolFrames>) + (size_t)flags_cap * 4 + (((size_t)rows_cap * 2 + 3) & ~(size_t)3);
  return (n + 15) & ~(size_t)15;
}

template <int kPoolFrames>
__device__ void process_pooled(const K1bArgs& a, PoolScratch<kPoolFrames>& ps, int lane) {
  __syncwarp();
  const int n = min(ps.n_cand, kCandCap);
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    if (i < n) {
      const uint32_t c = ps.cand[i];
      const int q = (int)(c >> 30);
      process_one_candidate<false>(ps.mv[q], a, ps.roi[q], ps.kl[q], (int)(c & 0x7fffu), (int)((c >> 15) & 0x7fffu));
    }
  }
  __syncwarp();
  if (lane == 0) ps.n_cand = 0;
  __syncwarp();
}

#ifndef MPE_K1B_POOL_WARPS
#define MPE_K1B_POOL_WARPS 4
#endif
#ifndef MPE_K1B_POOL_MINBLOCKS
#define MPE_K1B_POOL_MINBLOCKS 8
#endif
// Measured @8192 frames of 752x480, two batches in flight (the headline) / the kernel alone:
//   no pooling 5.59 M frames/s / 0.196 ms;  4 frames per warp, 4-warp CTAs, 64 registers 5.85 M / 0.213 ms;  3 frames 5.75 M / 0.169 ms;
//   2 frames 5.69 M / 0.150 ms;  4 frames with 2-warp CTAs 5.78 M, with 1-warp CTAs 5.64 M (64 registers) / 5.62 M (95 registers).
// Alone the kernel is a latency chain (four row scans, then one border-following pass), so fewer frames per warp finish sooner;
// next to the other batch's kernels what counts is the issue slots and registers it leaves them: 71 M instead of 133 M warp
// instructions.  The pipeline number decides.
constexpr int kPoolWarpsPerCta = MPE_K1B_POOL_WARPS;
template <int kPoolFrames>
__global__ void __launch_bounds__(32 * kPoolWarpsPerCta, MPE_K1B_POOL_MINBLOCKS) extract_blobs_pooled_kernel(const K1bArgs a, int flags_cap) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t k1b_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const K1Geom& g = a.g;
  const int f0 = (blockIdx.x * kPoolWarpsPerCta + warp) * kPoolFrames;
  if (f0 >= g.n_frames) return;
  const int rows_cap = g.mask_rows;
  uint8_t* my = k1b_smem + (size_t)warp * k1b_pool_stride<kPoolFrames>(flags_cap, rows_cap);
  PoolScratch<kPoolFrames>& ps = *reinterpret_cast<PoolScratch<kPoolFrames>*>(my);
  uint32_t* rowflags = reinterpret_cast<uint32_t*>(my + sizeof(PoolScratch<kPoolFrames>));
  uint16_t* rows = reinterpret_cast<uint16_t*>(my + sizeof(PoolScratch<kPoolFrames>) + (size_t)flags_cap * 4);
  if (lane == 0) ps.n_cand = 0;
  if (lane < kPoolFrames) { ps.kl[lane].n_kept = 0; ps.kl[lane].flags = 0; }
  __syncwarp();

  for (int q = 0; q < kPoolFrames; ++q) {
    const int f = f0 + q;
    if (f >= g.n_frames) break;                                // warp-uniform
    if (a.active && !a.active[f]) continue;
    const Roi roi = g.rois ? g.rois[f] : g.roi;
    const int roi_wpr = (roi.w + 31) >> 5;
    const int roi_n_ct = (roi.w + g.tw_px - 1) / g.tw_px;
    const int roi_strips = (roi.h + kTileRows - 1) / kTileRows;
    const uint32_t* gflags = a.rowflags + (size_t)f * g.flags_per_frame;
    if (lane == 0) ps.n_rows = 0;
    for (int i = lane; i < roi_strips * g.n_ct && i < flags_cap; i += 32) rowflags[i] = gflags[i];
    __syncwarp();
    list_rows(rowflags, roi_strips, g.n_ct, roi_n_ct, &ps.n_rows, rows, rows_cap, lane);
    __syncwarp();

    MaskView m;
    m.flags = rowflags;
    m.mask = a.mask + (size_t)f * g.mask_rows * g.mask_wpr;
    m.w = roi.w; m.h = roi.h;
    m.n_ct = g.n_ct;
    m.roi_n_ct = roi_n_ct;
    m.tw_px = g.tw_px;
    m.wpr = g.mask_wpr;
    m.words_per_ct = g.tw_px >> 5;
    m.win = nullptr; m.win_wpr = 0;
    if (lane == 0) {
      ps.mv[q] = m;
      ps.mv[q].flags = gflags;                                 // what the border follower of this frame reads later
      ps.roi[q] = roi;
    }
    __syncwarp();

    const int n_rows = min(ps.n_rows, rows_cap);
    for (int base = 0; base < n_rows; base += 32) {
      if (base + lane < n_rows) {
        const int y = rows[base + lane];
        unsigned long long nz = 0;
        for (int w0 = 0; w0 < roi_wpr; w0 += 8) {              // eight loads in flight
          uint32_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (w0 + u < roi_wpr) ? mv_word<false>(m, y, w0 + u) : 0u;
#pragma unroll
          for (int u = 0; u < 8; ++u) nz |= (unsigned long long)(v[u] != 0u) << ((w0 + u) & 63);
        }
        for (int wi = 0; wi < roi_wpr; ++wi) {
          if (roi_wpr <= 64 && !((nz >> wi) & 1ull)) continue;
          uint32_t cand = candidate_bits<false>(m, y, wi);
          while (cand) {
            const int bbit = __ffs(cand) - 1;
            cand &= cand - 1;
            const int slot = atomicAdd(&ps.n_cand, 1);
            if (slot < kCandCap) ps.cand[slot] = ((uint32_t)q << 30) | ((uint32_t)y << 15) | (uint32_t)(wi * 32 + bbit);
            else atomicOr(&ps.kl[q].flags, MPE_F_BLOB_OVERFLOW);
          }
        }
      }
      __syncwarp();
      if (ps.n_cand >= kCandCap / 2) process_pooled<kPoolFrames>(a, ps, lane);
    }
    __syncwarp();                                              // rowflags / rows are rewritten by the next frame
  }
  if (ps.n_cand > 0) process_pooled<kPoolFrames>(a, ps, lane);
  __syncwarp();

  // ---- order + undistortion over the pooled blobs, one per lane
  int offs[kPoolFrames + 1];
  offs[0] = 0;
#pragma unroll
  for (int q = 0; q < kPoolFrames; ++q) offs[q + 1] = offs[q] + min(ps.kl[q].n_kept, MPE_MAX_BLOBS);
  for (int i = lane; i < offs[kPoolFrames]; i += 32) {
    int q = 0;
#pragma unroll
    for (int t = 1; t < kPoolFrames; ++t) q += (i >= offs[t]);
    emit_blob(a, ps.kl[q], f0 + q, i - offs[q], offs[q + 1] - offs[q]);
  }
  if (lane < kPoolFrames) {
    const int f = f0 + lane;
    if (f < g.n_frames && !(a.active && !a.active[f])) {
      a.n_det[f] = offs[lane + 1] - offs[lane];
      a.flags[f] = ps.kl[lane].flags;
    }
  }
}

template <int kPoolFrames>
static cudaError_t launch_extract_blobs_pooled(const K1bArgs& a, cudaStream_t st) {
  static SmemAttrCache configured_pool;
  int flags_cap = a.g.n_strips * a.g.n_ct;                    // the row flags this launch's geometry can produce
  if (flags_cap > kMaxFlagWords) flags_cap = kMaxFlagWords;
  const size_t smem = k1b_pool_stride<kPoolFrames>(flags_cap, a.g.mask_rows) * kPoolWarpsPerCta;
  cudaError_t e = ensure_dynamic_smem(extract_blobs_pooled_kernel<kPoolFrames>, smem, configured_pool);
  if (e != cudaSuccess) return e;
  const int groups = (a.g.n_frames + kPoolFrames - 1) / kPoolFrames;
  return launch_k(extract_blobs_pooled_kernel<kPoolFrames>, (groups + kPoolWarpsPerCta - 1) / kPoolWarpsPerCta, 32 * kPoolWarpsPerCta, smem, st, a, flags_cap);
}

cudaError_t launch_extract_blobs(const K1bArgs& a, cudaStream_t st) {
  // frames per warp: MPE_K1B_POOL = 0 (none) / 2 / 4; default 4 for whole-image batches (the cold pipeline keeps two batches in
  // flight), none for per-frame ROIs: a tracking step runs alone on its stream, where the shorter chain wins (measured, 8192
  // streams: 7.38 M frames/s without, 7.36 M with two, 7.11 M with four frames per warp)
  static int pool_env = -1;
  if (pool_env < 0) { const char* e = getenv("MPE_K1B_POOL"); pool_env = e ? atoi(e) : -1; if (pool_env < -1) pool_env = -1; }
  int pool = (pool_env >= 0) ? pool_env : (a.g.rois ? 0 : 4);
  if (a.g.n_frames < kK1bPoolMinFrames || a.g.max_roi_w > 32767 || a.g.max_roi_h > 32767) pool = 0;
  if (pool >= 3) return launch_extract_blobs_pooled<4>(a, st);
  if (pool >= 1) return launch_extract_blobs_pooled<2>(a, st);
  int grid = (a.g.n_frames + kBlobWarpsPerCta - 1) / kBlobWarpsPerCta;
  static SmemAttrCache configured;
  const int flags_cap = a.g.flags_per_frame < kMaxFlagWords ? a.g.flags_per_frame : kMaxFlagWords;
  const int win_words = (a.g.n_frames <= kK1bWindowMaxFrames) ? kK1bWindowWords : 0;
  size_t smem = k1b_scratch_stride(flags_cap, a.g.mask_rows, win_words) * kBlobWarpsPerCta;
  {
    cudaError_t e = ensure_dynamic_smem(extract_blobs_kernel, smem, configured);
    if (e != cudaSuccess) return e;
  }
  return launch_k(extract_blobs_kernel, grid, 32 * kBlobWarpsPerCta, smem, st, a, win_words);
}

}  // namespace mpe
