"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/libpose_oracle.so (C++ restatement of the pose path) and
the estimateBodyPose state machine (/root/reference/.../src/pose_estimator.cpp:62-147) stitched from the
cv2 findLeds oracle and the C++ stage functions."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from . import find_leds_cv2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpose_oracle.so")
    src = os.path.join(_HERE, "pose_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libpose_oracle.so"], stdout=subprocess.DEVNULL)
    return so


dp, up, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint), C.POINTER(C.c_int)
# (name, argtypes, restype) of the C API; oracle/ref_pose.py binds the mper_* twins of the reference build with the same table
SIGNATURES = [
    ("mpeo_destroy", [C.c_void_p], None),
    ("mpeo_solve_quartic", [dp, dp], C.c_int),
    ("mpeo_p3p", [dp, dp, dp], C.c_int),
    ("mpeo_set_camera", [C.c_void_p, dp, dp, C.c_int], None),
    ("mpeo_set_markers", [C.c_void_p, dp, C.c_int], None),
    ("mpeo_set_params", [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double], None),
    ("mpeo_set_histogram_threshold", [C.c_void_p, C.c_uint], None),
    ("mpeo_get_histogram_threshold", [C.c_void_p], C.c_uint),
    ("mpeo_set_image_points", [C.c_void_p, dp, C.c_int], None),
    ("mpeo_get_image_vectors", [C.c_void_p, dp], C.c_int),
    ("mpeo_initialise", [C.c_void_p], C.c_uint),
    ("mpeo_get_histogram", [C.c_void_p, up], C.c_int),
    ("mpeo_get_counters", [C.c_void_p, C.POINTER(C.c_ulonglong)], None),
    ("mpeo_get_correspondences", [C.c_void_p, up], C.c_int),
    ("mpeo_set_correspondences", [C.c_void_p, up, C.c_int], None),
    ("mpeo_check_correspondences", [C.c_void_p], C.c_uint),
    ("mpeo_find_correspondences", [C.c_void_p], None),
    ("mpeo_optimise_pose", [C.c_void_p], C.c_int),
    ("mpeo_optimise_and_update_pose", [C.c_void_p], None),
    ("mpeo_last_gn_iterations", [C.c_void_p], C.c_int),
    ("mpeo_update_pose", [C.c_void_p], None),
    ("mpeo_get_predicted_pose", [C.c_void_p, dp], None),
    ("mpeo_set_predicted_pose", [C.c_void_p, dp, C.c_double], None),
    ("mpeo_get_current_pose", [C.c_void_p, dp], None),
    ("mpeo_get_previous_pose", [C.c_void_p, dp], None),
    ("mpeo_set_state", [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_uint], None),
    ("mpeo_get_covariance", [C.c_void_p, dp], None),
    ("mpeo_set_predicted_time", [C.c_void_p, C.c_double], None),
    ("mpeo_get_predicted_time", [C.c_void_p], C.c_double),
    ("mpeo_it_since_initialized", [C.c_void_p], C.c_uint),
    ("mpeo_predict_pose", [C.c_void_p, C.c_double], None),
    ("mpeo_predict_marker_positions", [C.c_void_p], None),
    ("mpeo_get_predicted_pixels", [C.c_void_p, dp], C.c_int),
    ("mpeo_set_predicted_pixels", [C.c_void_p, dp, C.c_int], None),
    ("mpeo_determine_roi", [C.c_void_p, C.c_int, C.c_int, C.c_int, ip], None),
    ("mpeo_project2d", [C.c_void_p, dp, dp, dp], None),
    ("mpeo_exponential_map", [dp, dp], None),
    ("mpeo_logarithm_map", [dp, dp], None),
    ("mpeo_distort_point", [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_float)], None),
    ("mpeo_cold_pose_batch", [C.c_void_p, dp, ip, C.c_int, C.c_int, dp, ip], C.c_int),
        ]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.mpeo_create.restype = C.c_void_p
        for name, args, res in SIGNATURES:
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _LIB = L
    return _LIB


_CLIB = None


class _Counted:
    """libpose_oracle_counted.so (the restatement compiled on a counting double, oracle/counted_double.h) under the mpeo_* names."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        if name.startswith("mpeo_"):
            return getattr(self._L, "mpeoc_" + name[len("mpeo_"):])
        return getattr(self._L, name)


def counted_lib():
    global _CLIB
    if _CLIB is None:
        so = os.path.join(_HERE, "libpose_oracle_counted.so")
        src = [os.path.join(_HERE, n) for n in ("pose_oracle.cpp", "pose_oracle_counted.cpp", "counted_double.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in src):
            subprocess.check_call(["make", "-C", _HERE, "libpose_oracle_counted.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.mpeoc_create.restype = C.c_void_p
        for name, args, res in SIGNATURES:
            fn = getattr(L, "mpeoc_" + name[len("mpeo_"):])
            fn.argtypes = args
            fn.restype = res
        L.mpeoc_ops_get.argtypes = [C.POINTER(C.c_ulonglong)]
        _CLIB = L
    return _CLIB


OP_CLASSES = ("add", "mul", "div", "sqrt", "cmp", "special")


def count_ops(K, D, markers, params, detections):
    """Exact FP64 operation counts of the CPU restatement for one frame's pose path, per stage:
    {'initialise': {...}, 'optimise': {...}} with the classes of OP_CLASSES plus 'flops' = add + mul + div + sqrt + cmp
    (transcendental library calls are listed separately as 'special').  `initialise` = setImagePoints + the brute-force
    sweep + histogram decode + checkCorrespondences (pose_estimator.cpp:544-721), `optimise` = optimisePose (:733-792)."""
    L = counted_lib()
    est = PoseEstimatorOracle(K, D, markers, params, _lib=_Counted(L))

    def read():
        out = (C.c_ulonglong * 6)()
        L.mpeoc_ops_get(out)
        d = dict(zip(OP_CLASSES, [int(v) for v in out]))
        d["flops"] = d["add"] + d["mul"] + d["div"] + d["sqrt"] + d["cmp"]
        return d

    L.mpeoc_ops_reset()
    est.set_image_points(detections)
    ok = est.initialise()
    init = read()
    opt = None
    if ok:
        L.mpeoc_ops_reset()
        iters = est.optimise_pose()
        opt = read()
        opt["gn_iterations"] = iters
    return {"ok": ok, "initialise": init, "optimise": opt, "histogram": est.histogram(), "correspondences": est.correspondences()}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def solve_quartic(factors):
    f = np.ascontiguousarray(factors, np.float64)
    r = np.zeros(4)
    lib().mpeo_solve_quartic(_dp(f), _dp(r))
    return r


def p3p(feature_vectors, world_points):
    """feature_vectors, world_points: 3x3 with one vector per COLUMN (as the reference).  Returns (rc, 4x3x4)."""
    f = np.ascontiguousarray(np.asarray(feature_vectors, np.float64).T)   # column k contiguous
    P = np.ascontiguousarray(np.asarray(world_points, np.float64).T)
    sol = np.zeros((4, 3, 4))
    rc = lib().mpeo_p3p(_dp(f), _dp(P), _dp(sol))
    return rc, sol


class PoseEstimatorOracle:
    """Mirror of monocular_pose_estimator::PoseEstimator (pose_estimator.h:52-803) on the CPU oracle."""
    MIN_NUM_LEDS_DETECTED = 4   # pose_estimator.h:78

    def __init__(self, K, D, markers, params, _lib=None):
        self.L = _lib if _lib is not None else lib()
        self.h = C.c_void_p(self.L.mpeo_create())
        self.K = np.ascontiguousarray(K, np.float64)
        self.D = np.ascontiguousarray(D, np.float64)
        self.params = params
        self.n_obj = len(markers)
        self.L.mpeo_set_camera(self.h, _dp(self.K), _dp(self.D), len(self.D))
        m = np.ascontiguousarray(markers, np.float64)
        self.L.mpeo_set_markers(self.h, _dp(m), len(m))
        self.L.mpeo_set_params(self.h, params.back_projection_pixel_tolerance, params.nearest_neighbour_pixel_tolerance,
                               params.certainty_threshold, params.valid_correspondence_threshold)
        self.region_of_interest = None
        self.distorted_detection_centers = np.zeros((0, 2), np.float32)
        self.pose_updated = False
        self.n_det = 0

    def __del__(self):
        try:
            self.L.mpeo_destroy(self.h)
        except Exception:
            pass

    # ---- thin accessors
    def set_image_points(self, pts):
        p = np.ascontiguousarray(pts, np.float64).reshape(-1, 2)
        self.n_det = len(p)
        self.L.mpeo_set_image_points(self.h, _dp(p), len(p))

    def image_vectors(self):
        out = np.zeros((max(self.n_det, 1), 3))
        n = self.L.mpeo_get_image_vectors(self.h, _dp(out))
        return out[:n]

    def initialise(self):
        return int(self.L.mpeo_initialise(self.h))

    def histogram(self):
        out = np.zeros(self.n_det * self.n_obj, np.uint32)
        n = self.L.mpeo_get_histogram(self.h, out.ctypes.data_as(C.POINTER(C.c_uint)))
        return out[:n].reshape(self.n_det, self.n_obj)

    def counters(self):
        out = (C.c_ulonglong * 3)()
        self.L.mpeo_get_counters(self.h, out)
        return dict(p3p=out[0], finite=out[1], voting=out[2])

    def correspondences(self):
        out = np.zeros((64, 2), np.uint32)
        n = self.L.mpeo_get_correspondences(self.h, out.ctypes.data_as(C.POINTER(C.c_uint)))
        return out[:n].copy()

    def set_correspondences(self, corr):
        c = np.ascontiguousarray(corr, np.uint32).reshape(-1, 2)
        self.L.mpeo_set_correspondences(self.h, c.ctypes.data_as(C.POINTER(C.c_uint)), len(c))

    def check_correspondences(self):
        return int(self.L.mpeo_check_correspondences(self.h))

    def find_correspondences(self):
        self.L.mpeo_find_correspondences(self.h)

    def optimise_pose(self):
        return int(self.L.mpeo_optimise_pose(self.h))

    def predicted_pose(self):
        out = np.zeros((4, 4)); self.L.mpeo_get_predicted_pose(self.h, _dp(out)); return out

    def set_predicted_pose(self, T, time=0.0):
        T = np.ascontiguousarray(T, np.float64); self.L.mpeo_set_predicted_pose(self.h, _dp(T), time)

    def current_pose(self):
        out = np.zeros((4, 4)); self.L.mpeo_get_current_pose(self.h, _dp(out)); return out

    def covariance(self):
        out = np.zeros((6, 6)); self.L.mpeo_get_covariance(self.h, _dp(out)); return out

    def gn_iterations(self):
        return int(self.L.mpeo_last_gn_iterations(self.h))

    def it_since_initialized(self):
        return int(self.L.mpeo_it_since_initialized(self.h))

    def predicted_pixels(self):
        out = np.zeros((self.n_obj, 2)); n = self.L.mpeo_get_predicted_pixels(self.h, _dp(out)); return out[:n]

    def determine_roi(self, width, height):
        roi = (C.c_int * 4)()
        self.L.mpeo_determine_roi(self.h, width, height, int(self.params.roi_border_thickness), roi)
        return tuple(roi)

    # ---- pose_estimator.cpp:62-147
    def _find_leds(self, image, roi):
        p = self.params
        px, centers = find_leds_cv2.find_leds(image, roi, p.threshold_value, p.gaussian_sigma, p.min_blob_area,
                                              p.max_blob_area, p.max_width_height_distortion,
                                              p.max_circular_distortion, self.K, self.D)
        self.distorted_detection_centers = centers
        return px

    def estimate_body_pose(self, image, time_to_predict):
        self.pose_updated = False
        h, w = image.shape
        detected = np.zeros((0, 2))
        if self.it_since_initialized() < 1:                                   # :68
            self.L.mpeo_set_predicted_time(self.h, time_to_predict)
            self.region_of_interest = (0, 0, w, h)
            px = self._find_leds(image, self.region_of_interest)
            if px is not None:
                detected = px
            if len(detected) >= self.MIN_NUM_LEDS_DETECTED:                    # :80
                self.set_image_points(detected)
                if self.initialise() == 1:
                    self._optimise_and_update_pose()
        else:                                                                  # :97
            self._predict_with_roi(time_to_predict, w, h)
            px = self._find_leds(image, self.region_of_interest)
            if px is not None:
                detected = px
            num_loops = 0
            while True:
                num_loops += 1
                if len(detected) >= self.MIN_NUM_LEDS_DETECTED:
                    self.set_image_points(detected)
                    self._find_correspondences_and_predict_pose()
                    break
                if num_loops < 2:
                    self.region_of_interest = (0, 0, w, h)
                    px = self._find_leds(image, self.region_of_interest)
                    if px is not None:         # untouched when nothing found (led_detector.cpp:91)
                        detected = px
                else:
                    break
        return self.pose_updated

    def _optimise_and_update_pose(self):                                       # :802-812
        self.L.mpeo_optimise_and_update_pose(self.h)
        self.pose_updated = True

    def _predict_with_roi(self, t, w, h):                                      # :814-829
        if self.it_since_initialized() >= 2:
            self.L.mpeo_predict_pose(self.h, t)
        else:
            self.L.mpeo_set_predicted_time(self.h, t)
        self.L.mpeo_predict_marker_positions(self.h)
        self.region_of_interest = self.determine_roi(w, h)

    def _find_correspondences_and_predict_pose(self):                          # :831-848
        self.find_correspondences()
        if self.check_correspondences() == 1:
            self._optimise_and_update_pose()
        elif self.initialise() == 1:
            self._optimise_and_update_pose()
