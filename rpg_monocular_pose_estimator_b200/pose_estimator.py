"""Host-side mirror of monocular_pose_estimator::PoseEstimator
(/root/reference/monocular_pose_estimator_lib/include/monocular_pose_estimator_lib/pose_estimator.h:52-803,
src/pose_estimator.cpp) on top of the C ABI.  Method names follow the reference (snake_case aliases are not
provided on purpose: tests read like the reference's call sites).  The per-frame state machine
(estimateBodyPose, pose_estimator.cpp:62-147) and the tiny sequential helpers of tracking mode (predictPose,
determineROI, findCorrespondences) run on the host exactly as in the reference; every heavy stage is a CUDA
kernel behind the ABI.
"""
from __future__ import annotations

import ctypes as C
import math
import numpy as np

from . import _lib
from ._lib import MpeError, MpeParams, MpeRect, MpeResult, MPE_MAX_BLOBS, MPE_MAX_DET, MPE_MAX_LEDS
from .synth import Params


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """Owns one mpe_ctx (one per GPU, single caller)."""

    def __init__(self, device: int = 0, max_batch: int = 1, max_width: int = 752, max_height: int = 480):
        self.L = _lib.load_library()
        h = C.c_void_p()
        rc = self.L.mpe_create(C.byref(h), device, max_batch, max_width, max_height)
        if rc != 0:
            raise MpeError(f"mpe_create failed ({rc}) — is a CUDA device visible? There is no CPU fallback.")
        self.h = h
        self.max_batch, self.max_width, self.max_height = max_batch, max_width, max_height
        self.n_obj = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.mpe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise MpeError(f"mpe error {rc}: {self.L.mpe_last_error(self.h).decode()}")

    def set_camera(self, K, D):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        D = np.ascontiguousarray(D, np.float64).reshape(-1)
        self.check(self.L.mpe_set_camera(self.h, _dp(K), _dp(D), len(D)))

    def set_markers(self, xyz):
        m = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        self.check(self.L.mpe_set_markers(self.h, _dp(m), len(m)))
        self.n_obj = len(m)

    def set_params(self, p: Params):
        s = MpeParams(int(p.threshold_value), int(p.roi_border_thickness), p.gaussian_sigma, p.min_blob_area, p.max_blob_area,
                      p.max_width_height_distortion, p.max_circular_distortion, p.back_projection_pixel_tolerance,
                      p.nearest_neighbour_pixel_tolerance, p.certainty_threshold, p.valid_correspondence_threshold)
        self.check(self.L.mpe_set_params(self.h, C.byref(s)))

    def set_stream(self, cuda_stream_ptr: int):
        self.check(self.L.mpe_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # ---- stage calls
    def find_leds(self, image: np.ndarray, roi):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        px = np.zeros((MPE_MAX_BLOBS, 2), np.float64)
        ce = np.zeros((MPE_MAX_BLOBS, 2), np.float32)
        n, fl = C.c_int(0), C.c_int(0)
        r = MpeRect(int(roi[0]), int(roi[1]), int(roi[2]), int(roi[3]))
        self.check(self.L.mpe_find_leds(self.h, img.ctypes.data_as(C.c_void_p), img.strides[0], w, h, r, _dp(px),
                                        ce.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n), C.byref(fl)))
        return px[:n.value].copy(), ce[:n.value].copy(), fl.value

    def find_leds_strided(self, image: np.ndarray, roi):
        """Like find_leds but takes a row-strided (non-contiguous) uint8 view as is: pitch = image.strides[0]."""
        assert image.dtype == np.uint8 and image.strides[1] == 1
        h, w = image.shape
        px = np.zeros((MPE_MAX_BLOBS, 2), np.float64)
        ce = np.zeros((MPE_MAX_BLOBS, 2), np.float32)
        n, fl = C.c_int(0), C.c_int(0)
        r = MpeRect(int(roi[0]), int(roi[1]), int(roi[2]), int(roi[3]))
        self.check(self.L.mpe_find_leds(self.h, C.c_void_p(image.ctypes.data), image.strides[0], w, h, r, _dp(px),
                                        ce.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n), C.byref(fl)))
        return px[:n.value].copy(), ce[:n.value].copy(), fl.value

    def initialise(self, det):
        det = np.ascontiguousarray(det, np.float64).reshape(-1, 2)
        hist = np.zeros((len(det), self.n_obj), np.uint32)
        corr = np.zeros((MPE_MAX_LEDS, 2), np.uint32)
        k, ok = C.c_int(0), C.c_int(0)
        pose = np.zeros((4, 4))
        self.check(self.L.mpe_initialise(self.h, _dp(det), len(det), hist.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         corr.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(k), _dp(pose), C.byref(ok)))
        return ok.value, hist, corr[:k.value].copy(), pose

    def check_correspondences(self, det, corr):
        det = np.ascontiguousarray(det, np.float64).reshape(-1, 2)
        corr = np.ascontiguousarray(corr, np.uint32).reshape(-1, 2)
        ok = C.c_int(0)
        pose = np.zeros((4, 4))
        self.check(self.L.mpe_check_correspondences(self.h, _dp(det), len(det), corr.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                    len(corr), _dp(pose), C.byref(ok)))
        return ok.value, pose

    def optimise_pose(self, det, corr, pose):
        det = np.ascontiguousarray(det, np.float64).reshape(-1, 2)
        corr = np.ascontiguousarray(corr, np.uint32).reshape(-1, 2)
        pose = np.ascontiguousarray(pose, np.float64).copy()
        cov = np.zeros((6, 6))
        it = C.c_int(0)
        self.check(self.L.mpe_optimise_pose(self.h, _dp(det), len(det), corr.ctypes.data_as(C.POINTER(C.c_uint32)), len(corr),
                                            _dp(pose), _dp(cov), C.byref(it)))
        return pose, cov, it.value

    def p3p(self, feature_vectors, world_points):
        """feature_vectors, world_points: (n,3,3) with one vector per COLUMN (reference layout)."""
        f = np.ascontiguousarray(np.transpose(np.asarray(feature_vectors, np.float64), (0, 2, 1)))
        P = np.ascontiguousarray(np.transpose(np.asarray(world_points, np.float64), (0, 2, 1)))
        n = len(f)
        sol = np.zeros((n, 4, 3, 4))
        st = np.zeros(n, np.int32)
        self.check(self.L.mpe_p3p_compute_poses(self.h, _dp(f), _dp(P), n, _dp(sol), st.ctypes.data_as(C.POINTER(C.c_int))))
        return st, sol

    # ---- batch
    def estimate_batch(self, frames: np.ndarray):
        """frames: (B,H,W) uint8 host array.  Cold mode (every frame uninitialised).  Returns ctypes array of MpeResult."""
        fr = np.ascontiguousarray(frames, np.uint8)
        B, H, W = fr.shape
        res = (MpeResult * B)()
        self.check(self.L.mpe_estimate_batch(self.h, fr.ctypes.data_as(C.c_void_p), fr.strides[1], fr.strides[0], W, H, B, res))
        return res

    def estimate_batch_device(self, dev_ptr: int, pitch: int, frame_stride: int, width: int, height: int, n: int):
        res = (MpeResult * n)()
        self.check(self.L.mpe_estimate_batch_device(self.h, C.c_void_p(dev_ptr), pitch, frame_stride, width, height, n, res))
        return res

    def estimate_batch_device_async(self, dev_ptr: int, pitch: int, frame_stride: int, width: int, height: int, n: int):
        self.check(self.L.mpe_estimate_batch_device_async(self.h, C.c_void_p(dev_ptr), pitch, frame_stride, width, height, n))

    def fetch_results(self, n: int):
        res = (MpeResult * n)()
        self.check(self.L.mpe_fetch_results(self.h, n, res))
        return res

    # ---- device-resident tracking loop (one PoseEstimator per stream, state on the GPU)
    def streams_reset(self, n_streams: int):
        self.check(self.L.mpe_streams_reset(self.h, n_streams))

    def streams_set_frame_map(self, frame_index_device_ptr: int, n_frames_in_buffer: int):
        self.check(self.L.mpe_streams_set_frame_map(self.h, C.c_void_p(frame_index_device_ptr), n_frames_in_buffer))

    def streams_step_device(self, dev_ptr: int, pitch: int, frame_stride: int, width: int, height: int, times, fetch: bool = True):
        """One estimateBodyPose step for len(times) streams; frame s of the device buffer belongs to stream s (or
        frame_map[s] when a frame map is set)."""
        t = np.ascontiguousarray(times, np.float64)
        n = len(t)
        res = (MpeResult * n)() if fetch else None
        self.check(self.L.mpe_streams_step_device(self.h, C.c_void_p(dev_ptr), pitch, frame_stride, width, height, n, _dp(t), res))
        return res

    def streams_step(self, frames: np.ndarray, times, out=None):
        """One estimateBodyPose step for len(times) streams from HOST images (S,H,W) uint8 (image s -> stream s): the per-image
        call of a camera driver.  Returns the ctypes MpeResult array (pass `out` to reuse one)."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        t = np.ascontiguousarray(times, np.float64)
        n, H, W = frames.shape
        assert len(t) == n
        res = out if out is not None else (MpeResult * n)()
        self.check(self.L.mpe_streams_step(self.h, C.c_void_p(frames.ctypes.data), frames.strides[1], frames.strides[0], W, H, n, _dp(t), res))
        return res

    INGEST_COPY, INGEST_ZERO_COPY, INGEST_AUTO = 0, 1, 2

    def set_ingest_mode(self, mode: int):
        self.check(self.L.mpe_set_ingest_mode(self.h, int(mode)))

    def ingest_stats(self):
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self.check(self.L.mpe_get_ingest_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"copy_steps": a.value, "zero_copy_steps": b.value, "h2d_bytes_copied": c.value}

    def set_k2_filter(self, mode):
        """0 / False: every hypothesis of the sweep is scored exactly; 1: reject filter behind the exact P3P solve;
        2 / True (default): tier-1 pre-test in front of it (csrc/p3p_tier1.cuh).  Results are identical in all modes."""
        m = 2 if mode is True else (0 if mode is False else int(mode))
        self.check(self.L.mpe_set_k2_filter(self.h, m))

    def set_graph_replay(self, on: bool):
        self.check(self.L.mpe_set_graph_replay(self.h, 1 if on else 0))

    def copy_poses_device(self, dst_device_ptr: int, n: int):
        self.check(self.L.mpe_copy_poses_device(self.h, n, C.c_void_p(dst_device_ptr)))

    def copy_results_device(self, dst_device_ptr: int, n: int):
        self.check(self.L.mpe_copy_results_device(self.h, n, C.c_void_p(dst_device_ptr)))

    def probe_fp64_peak(self) -> float:
        """Measured FP64 throughput of the device in TFLOP/s (FMA = 2 flops)."""
        v = C.c_double(0)
        self.check(self.L.mpe_probe_fp64_peak(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self.check(self.L.mpe_synchronize(self.h))

    def enable_kernel_timing(self, on=True):
        self.check(self.L.mpe_enable_kernel_timing(self.h, 1 if on else 0))

    def kernel_times_ms(self):
        out = (C.c_float * 5)()
        self.check(self.L.mpe_get_kernel_times(self.h, out))
        return list(out)

    def launch_count(self):
        return int(self.L.mpe_kernel_launch_count(self.h))


def results_to_arrays(res):
    """ctypes MpeResult array -> dict of numpy arrays."""
    n = len(res)
    buf = np.frombuffer(res, dtype=np.uint8).reshape(n, C.sizeof(MpeResult))
    dt = np.dtype([("updated", "i4"), ("n_det", "i4"), ("n_corr", "i4"), ("gn_iters", "i4"), ("flags", "i4"), ("init_ok", "i4"),
                   ("roi", "i4", 4), ("pose", "f8", 16), ("cov", "f8", 36), ("corr", "u4", 2 * MPE_MAX_LEDS),
                   ("det", "f8", 2 * MPE_MAX_DET), ("centers", "f4", 2 * MPE_MAX_DET)], align=True)
    assert dt.itemsize == C.sizeof(MpeResult), (dt.itemsize, C.sizeof(MpeResult))
    return buf.view(dt).reshape(n)


class PoseEstimator:
    """monocular_pose_estimator::PoseEstimator with the reference's public surface (pose_estimator.h:334-801)."""
    min_num_leds_detected_ = 4   # pose_estimator.h:78

    def __init__(self, context: Context | None = None, width: int = 752, height: int = 480):
        self.ctx = context or Context(0, 1, width, height)
        # public tunables (pose_estimator.h:82-91) — written directly by the caller, like MPENode does
        self.camera_matrix_K_ = None
        self.camera_distortion_coeffs_ = None
        self.detection_threshold_value_ = 140
        self.gaussian_sigma_ = 0.6
        self.min_blob_area_ = 10.0
        self.max_blob_area_ = 200.0
        self.max_width_height_distortion_ = 0.5
        self.max_circular_distortion_ = 0.5
        self.roi_border_thickness_ = 20
        # constructor defaults (pose_estimator.cpp:36-41)
        self.back_projection_pixel_tolerance_ = 3.0
        self.nearest_neighbour_pixel_tolerance_ = 5.0
        self.certainty_threshold_ = 0.75
        self.valid_correspondence_threshold_ = 0.7
        self.it_since_initialized_ = 0
        self.histogram_threshold_ = 0
        self.object_points_ = np.zeros((0, 4))
        self.image_points_ = np.zeros((0, 2))
        self.predicted_pixel_positions_ = np.zeros((0, 2))
        self.correspondences_ = np.zeros((0, 2), np.uint32)
        self.current_pose_ = np.eye(4)
        self.previous_pose_ = np.eye(4)
        self.predicted_pose_ = np.eye(4)
        self.pose_covariance_ = np.zeros((6, 6))
        self.current_time_ = 0.0
        self.previous_time_ = 0.0
        self.predicted_time_ = 0.0
        self.region_of_interest_ = (0, 0, width, height)
        self.distorted_detection_centers_ = np.zeros((0, 2), np.float32)
        self.pose_updated_ = False
        self.last_gn_iterations = 0
        self.last_flags = 0
        self._pushed = None

    # ---- configuration plumbing ------------------------------------------------------------------
    def configure(self, K, D, markers, params: Params):
        """Convenience: what MPENode's callbacks do (monocular_pose_estimator.cpp:84,110-120,222-233)."""
        self.camera_matrix_K_ = np.array(K, np.float64)
        self.camera_distortion_coeffs_ = np.array(D, np.float64)
        self.detection_threshold_value_ = params.threshold_value
        self.gaussian_sigma_ = params.gaussian_sigma
        self.min_blob_area_ = params.min_blob_area
        self.max_blob_area_ = params.max_blob_area
        self.max_width_height_distortion_ = params.max_width_height_distortion
        self.max_circular_distortion_ = params.max_circular_distortion
        self.roi_border_thickness_ = params.roi_border_thickness
        self.setBackProjectionPixelTolerance(params.back_projection_pixel_tolerance)
        self.setNearestNeighbourPixelTolerance(params.nearest_neighbour_pixel_tolerance)
        self.setCertaintyThreshold(params.certainty_threshold)
        self.setValidCorrespondenceThreshold(params.valid_correspondence_threshold)
        self.setMarkerPositions(np.hstack([np.asarray(markers, np.float64), np.ones((len(markers), 1))]))

    def _push(self):
        p = Params(int(self.detection_threshold_value_), float(self.gaussian_sigma_), float(self.min_blob_area_),
                   float(self.max_blob_area_), float(self.max_width_height_distortion_), float(self.max_circular_distortion_),
                   float(self.back_projection_pixel_tolerance_), float(self.nearest_neighbour_pixel_tolerance_),
                   float(self.certainty_threshold_), float(self.valid_correspondence_threshold_), int(self.roi_border_thickness_))
        key = (tuple(np.asarray(self.camera_matrix_K_).ravel()), tuple(np.asarray(self.camera_distortion_coeffs_).ravel()), p,
               self.histogram_threshold_)
        if key != self._pushed:
            self.ctx.set_camera(self.camera_matrix_K_, self.camera_distortion_coeffs_)
            self.ctx.set_params(p)
            self.ctx.check(self.ctx.L.mpe_set_histogram_threshold(self.ctx.h, int(self.histogram_threshold_)))
            self._pushed = key

    # ---- setters / getters (pose_estimator.cpp:50-60, 149-230, 246-286, 723-731) -------------------
    def setMarkerPositions(self, positions_of_markers_on_object):
        self.object_points_ = np.array(positions_of_markers_on_object, np.float64).reshape(-1, 4)
        self.predicted_pixel_positions_ = np.zeros((len(self.object_points_), 2))
        self.ctx.set_markers(self.object_points_[:, :3])
        self.histogram_threshold_ = int(self.ctx.L.mpe_get_histogram_threshold(self.ctx.h))   # numCombinations(n,3)
        self._pushed = None

    def getMarkerPositions(self): return self.object_points_
    def setPredictedPose(self, pose, time): self.predicted_pose_ = np.array(pose, np.float64); self.predicted_time_ = time
    def getPredictedPose(self): return self.predicted_pose_
    def getPoseCovariance(self): return self.pose_covariance_
    def setImagePoints(self, points): self.image_points_ = np.array(points, np.float64).reshape(-1, 2)
    def getImagePoints(self): return self.image_points_
    def setPredictedPixels(self, points): self.predicted_pixel_positions_ = np.array(points, np.float64).reshape(-1, 2)
    def getPredictedPixelPositions(self): return self.predicted_pixel_positions_
    def setCorrespondences(self, corrs): self.correspondences_ = np.array(corrs, np.uint32).reshape(-1, 2)
    def getCorrespondences(self): return self.correspondences_
    def setBackProjectionPixelTolerance(self, t): self.back_projection_pixel_tolerance_ = float(t)
    def getBackProjectionPixelTolerance(self): return self.back_projection_pixel_tolerance_
    def setNearestNeighbourPixelTolerance(self, t): self.nearest_neighbour_pixel_tolerance_ = float(t)
    def getNearestNeighbourPixelTolerance(self): return self.nearest_neighbour_pixel_tolerance_
    def setCertaintyThreshold(self, t): self.certainty_threshold_ = float(t)
    def getCertaintyThreshold(self): return self.certainty_threshold_
    def setValidCorrespondenceThreshold(self, t): self.valid_correspondence_threshold_ = float(t)
    def getValidCorrespondenceThreshold(self): return self.valid_correspondence_threshold_
    def setHistogramThreshold(self, t): self.histogram_threshold_ = int(t)
    def getHistogramThreshold(self): return self.histogram_threshold_
    def setPredictedTime(self, t): self.predicted_time_ = t
    def getPredictedTime(self): return self.predicted_time_

    # ---- small host math (sequential, a few hundred flops per frame): ONE implementation, in the library -------------
    # (csrc/tracking_math.cuh: the functions K4 runs per stream on the GPU, exported for the host as mpe_host_*)
    def project2d(self, point, transform):
        """pose_estimator.cpp:251-268"""
        K = np.ascontiguousarray(self.camera_matrix_K_, np.float64)
        T = np.ascontiguousarray(transform, np.float64)
        p = np.ascontiguousarray(np.asarray(point, np.float64)[:3])
        out = np.zeros(2)
        self.ctx.check(self.ctx.L.mpe_host_project_markers(_dp(K), _dp(T), _dp(p), 1, _dp(out)))
        return out

    def predictMarkerPositionsInImage(self):
        """pose_estimator.cpp:270-276"""
        K = np.ascontiguousarray(self.camera_matrix_K_, np.float64)
        T = np.ascontiguousarray(self.predicted_pose_, np.float64)
        m = np.ascontiguousarray(np.asarray(self.object_points_, np.float64)[:, :3])
        out = np.zeros((len(m), 2))
        self.ctx.check(self.ctx.L.mpe_host_project_markers(_dp(K), _dp(T), _dp(m), len(m), _dp(out)))
        self.predicted_pixel_positions_ = out

    @staticmethod
    def exponentialMap(twist):
        """pose_estimator.cpp:962-994"""
        from . import _lib
        t = np.ascontiguousarray(twist, np.float64); out = np.zeros((4, 4))
        _lib.load_library().mpe_host_exponential_map(_dp(t), _dp(out))
        return out

    @staticmethod
    def logarithmMap(trans):
        """pose_estimator.cpp:996-1064 (same special cases)"""
        from . import _lib
        T = np.ascontiguousarray(trans, np.float64); out = np.zeros(6)
        _lib.load_library().mpe_host_logarithm_map(_dp(T), _dp(out))
        return out

    def predictPose(self, time_to_predict):
        """pose_estimator.cpp:232-244"""
        self.predicted_time_ = time_to_predict
        out = np.zeros((4, 4))
        self.ctx.check(self.ctx.L.mpe_host_predict_pose(_dp(np.ascontiguousarray(self.previous_pose_, np.float64)),
                                                        _dp(np.ascontiguousarray(self.current_pose_, np.float64)),
                                                        float(self.previous_time_), float(self.current_time_), float(time_to_predict), _dp(out)))
        self.predicted_pose_ = out

    def findCorrespondences(self):
        """pose_estimator.cpp:372-392 with calculateMinDistancesAndPairs :862-906"""
        corr = []
        for i, p in enumerate(self.predicted_pixel_positions_):
            best, bj = math.inf, 0
            for j, q in enumerate(self.image_points_):
                d2 = float((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2)
                if d2 < best:
                    best, bj = d2, j + 1
            if math.sqrt(best) <= self.nearest_neighbour_pixel_tolerance_:
                corr.append((i + 1, bj))
        self.correspondences_ = np.array(corr, np.uint32).reshape(-1, 2)
        if len(self.image_points_) > MPE_MAX_DET:
            self._compact_to_matched_detections()

    def _compact_to_matched_detections(self):
        """Capacity path, same as the device loop (csrc/k4_tracking.cu): the brute-force tables hold MPE_MAX_DET detections, so with
        more than that the detections that are some LED's nearest neighbour are compacted to the front (ascending) and the
        correspondence rows renumbered; checkCorrespondences / optimisePose continue on that list.  last_flags gets MPE_F_TOO_MANY_DET."""
        used = sorted({int(c[1]) - 1 for c in self.correspondences_})
        remap = {j: m + 1 for m, j in enumerate(used)}
        self.image_points_ = np.asarray(self.image_points_, np.float64).reshape(-1, 2)[used].copy()
        if len(self.distorted_detection_centers_) > max(used, default=-1):
            self.distorted_detection_centers_ = np.asarray(self.distorted_detection_centers_)[used].copy()
        self.correspondences_ = np.array([(c[0], remap[int(c[1]) - 1]) for c in self.correspondences_], np.uint32).reshape(-1, 2)
        self.last_flags = getattr(self, "last_flags", 0) | 4          # MPE_F_TOO_MANY_DET

    # ---- device stages ----------------------------------------------------------------------------
    def initialise(self):
        """pose_estimator.cpp:544-721 (K2 sweep + decode, K3 check) -> 0/1.  More than MPE_MAX_DET (16) detections: the sweep's
        tables do not hold them — no exception, 0 is returned with MPE_F_TOO_MANY_DET in last_flags (limits)."""
        self._push()
        if len(self.image_points_) > MPE_MAX_DET:
            self.last_flags = getattr(self, "last_flags", 0) | 4
            return 0
        ok, hist, corr, pose = self.ctx.initialise(self.image_points_)
        self.last_histogram = hist
        self.correspondences_ = corr
        if ok:
            self.predicted_pose_ = pose
        return ok

    def checkCorrespondences(self):
        """pose_estimator.cpp:394-542 -> 0/1"""
        self._push()
        if len(self.correspondences_) < 4:
            return 0
        if len(self.image_points_) > MPE_MAX_DET:
            self._compact_to_matched_detections()
            if len(self.image_points_) < 4:
                return 0
        ok, pose = self.ctx.check_correspondences(self.image_points_, self.correspondences_)
        if ok:
            self.predicted_pose_ = pose
        return ok

    def optimisePose(self):
        """pose_estimator.cpp:733-792"""
        self._push()
        pose, cov, it = self.ctx.optimise_pose(self.image_points_, self.correspondences_, self.predicted_pose_)
        self.predicted_pose_, self.pose_covariance_, self.last_gn_iterations = pose, cov, it

    def updatePose(self):
        """pose_estimator.cpp:794-800"""
        self.previous_pose_ = self.current_pose_; self.current_pose_ = self.predicted_pose_
        self.previous_time_ = self.current_time_; self.current_time_ = self.predicted_time_

    def optimiseAndUpdatePose(self, time_to_predict):
        """pose_estimator.cpp:802-812"""
        self.optimisePose()
        if self.it_since_initialized_ < 2:
            self.it_since_initialized_ += 1
        self.updatePose()
        self.pose_updated_ = True

    def predictWithROI(self, time_to_predict, image):
        """pose_estimator.cpp:814-829"""
        from .led_detector import LEDDetector
        if self.it_since_initialized_ >= 2:
            self.predictPose(time_to_predict)
        else:
            self.setPredictedTime(time_to_predict)
        self.predictMarkerPositionsInImage()
        h, w = image.shape
        self.region_of_interest_ = LEDDetector.determineROI(self.getPredictedPixelPositions(), (w, h), self.roi_border_thickness_,
                                                            self.camera_matrix_K_, self.camera_distortion_coeffs_)

    def findCorrespondencesAndPredictPose(self, time_to_predict):
        """pose_estimator.cpp:831-848"""
        self.findCorrespondences()
        if self.checkCorrespondences() == 1:
            self.optimiseAndUpdatePose(time_to_predict)
        elif self.initialise() == 1:
            self.optimiseAndUpdatePose(time_to_predict)

    def _findLeds(self, image):
        from .led_detector import LEDDetector
        self._push()
        px, centers, flags = LEDDetector.findLeds(image, self.region_of_interest_, context=self.ctx)
        self.distorted_detection_centers_ = centers
        self.last_flags = flags
        return px

    def augmentImage(self, image):
        """pose_estimator.cpp:44-48: draws the body axes, the detections and the region of interest into a 3-channel image (host
        side, OpenCV drawing — as in the reference this is debug output next to the pose path, not part of it)."""
        from .visualization import Visualization
        return Visualization.createVisualizationImage(image, self.predicted_pose_, self.camera_matrix_K_, self.camera_distortion_coeffs_,
                                                      self.region_of_interest_, self.distorted_detection_centers_)

    def estimateBodyPose(self, image, time_to_predict):
        """pose_estimator.cpp:62-147"""
        self.pose_updated_ = False
        h, w = image.shape
        detected = np.zeros((0, 2))
        if self.it_since_initialized_ < 1:
            self.setPredictedTime(time_to_predict)
            self.region_of_interest_ = (0, 0, w, h)
            px = self._findLeds(image)
            if len(px) > 0:
                detected = px
            if len(detected) >= self.min_num_leds_detected_:
                self.setImagePoints(detected)
                if self.initialise() == 1:
                    self.optimiseAndUpdatePose(time_to_predict)
        else:
            self.predictWithROI(time_to_predict, image)
            px = self._findLeds(image)
            if len(px) > 0:
                detected = px
            num_loops = 0
            while True:
                num_loops += 1
                if len(detected) >= self.min_num_leds_detected_:
                    self.setImagePoints(detected)
                    self.findCorrespondencesAndPredictPose(time_to_predict)
                    break
                if num_loops < 2:
                    self.region_of_interest_ = (0, 0, w, h)
                    px = self._findLeds(image)
                    if len(px) > 0:          # pixel_positions untouched when nothing found (led_detector.cpp:91)
                        detected = px
                else:
                    break
        return self.pose_updated_
