"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/_ref/libref_pose.so: the UNMODIFIED reference sources
(/root/reference/monocular_pose_estimator_lib/src/{p3p,combinations,pose_estimator,led_detector}.cpp) compiled against the
Eigen / OpenCV stand-ins under oracle/eigen_shim and oracle/cv_shim (oracle/Makefile, target `ref`).

The seven OpenCV calls of LEDDetector::findLeds (led_detector.cpp:44,51,57,67,68,72,97) are bound here to the cv2 4.13
functions of the same name through C callbacks, so the reference's own findLeds / estimateBodyPose source runs on the real
OpenCV kernels.  Used by tests/ and by the fixture generator tests/golden/make_ref_golden.py only; it needs /root/reference
to build and therefore never runs on the GPU box (the prebuilt .so travels, tests skip when it is absent).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import cv2
import numpy as np

from . import pose_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_pose.so")
_LIB = None
_KEEP = []          # callback objects and the arrays their out-pointers refer to

ucp, ip, dp, fp, up = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float),
                       C.POINTER(C.c_uint))


class _Callbacks(C.Structure):
    _fields_ = [
        ("threshold", C.CFUNCTYPE(None, ucp, C.c_int, C.c_int, C.c_long, C.c_double, C.c_double, C.c_int, ucp)),
        ("gaussian_blur", C.CFUNCTYPE(None, ucp, C.c_int, C.c_int, C.c_long, C.c_double, C.c_double, C.c_int, ucp)),
        ("find_contours", C.CFUNCTYPE(C.c_int, ucp, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.POINTER(ip), C.POINTER(ip))),
        ("contour_area", C.CFUNCTYPE(C.c_double, ip, C.c_int)),
        ("bounding_rect", C.CFUNCTYPE(None, ip, C.c_int, ip)),
        ("moments", C.CFUNCTYPE(None, ip, C.c_int, dp)),
        ("undistort_points", C.CFUNCTYPE(None, fp, C.c_int, dp, dp, C.c_int, dp, fp)),
        ("project_points", C.CFUNCTYPE(None, fp, C.c_int, dp, dp, dp, dp, C.c_int, fp)),
        ("draw", C.CFUNCTYPE(None, ucp, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, ip, dp, C.c_int)),
    ]


def _img(ptr, rows, cols, step):
    buf = np.ctypeslib.as_array(ptr, shape=(rows * step,)) if rows * step > 0 else np.zeros(0, np.uint8)
    return np.lib.stride_tricks.as_strided(buf, shape=(rows, cols), strides=(step, 1))


def _contour(pts, n):
    return np.ctypeslib.as_array(pts, shape=(n, 2)).astype(np.int32).reshape(n, 1, 2)


def _make_callbacks():
    state = {}

    def threshold(src, rows, cols, step, thresh, maxval, typ, dst):
        _, out = cv2.threshold(_img(src, rows, cols, step), thresh, maxval, typ)
        np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = out

    def gaussian_blur(src, rows, cols, step, sx, sy, border, dst):
        out = cv2.GaussianBlur(np.ascontiguousarray(_img(src, rows, cols, step)), (0, 0), sx, sigmaY=sy, borderType=border)
        np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = out

    def find_contours(img, rows, cols, step, mode, method, counts_out, points_out):
        contours, _ = cv2.findContours(np.ascontiguousarray(_img(img, rows, cols, step)), mode, method)
        counts = np.array([len(c) for c in contours], np.int32)
        pts = (np.concatenate([c.reshape(-1, 2) for c in contours]).astype(np.int32) if len(contours)
               else np.zeros((0, 2), np.int32))
        pts = np.ascontiguousarray(pts)
        state["contours"] = (counts, pts)               # keep alive until the next call
        counts_out[0] = counts.ctypes.data_as(ip)
        points_out[0] = pts.ctypes.data_as(ip)
        return len(contours)

    def contour_area(pts, n):
        return float(cv2.contourArea(_contour(pts, n)))

    def bounding_rect(pts, n, out):
        x, y, w, h = cv2.boundingRect(_contour(pts, n))
        out[0], out[1], out[2], out[3] = x, y, w, h

    def moments(pts, n, out):
        m = cv2.moments(_contour(pts, n), False)
        for i, k in enumerate(("m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03")):
            out[i] = m[k]

    def undistort_points(src, n, K, D, nD, P, dst):
        s = np.ctypeslib.as_array(src, shape=(n, 2)).astype(np.float32).reshape(n, 1, 2)
        Km = np.ctypeslib.as_array(K, shape=(3, 3)).copy()
        Pm = np.ctypeslib.as_array(P, shape=(3, 3)).copy()
        Dv = np.ctypeslib.as_array(D, shape=(nD,)).copy() if nD > 0 else np.zeros(0)
        out = cv2.undistortPoints(s, Km, Dv, None, Pm)
        np.ctypeslib.as_array(dst, shape=(n, 2))[:] = out.reshape(n, 2)

    def project_points(xyz, n, rvec, tvec, K, D, nD, out):
        pts = np.ctypeslib.as_array(xyz, shape=(n, 3)).astype(np.float32).reshape(n, 1, 3)
        r = np.ctypeslib.as_array(rvec, shape=(3,)).copy().reshape(3, 1)
        t = np.ctypeslib.as_array(tvec, shape=(3,)).copy().reshape(3, 1)
        Km = np.ctypeslib.as_array(K, shape=(3, 3)).copy()
        Dv = np.ctypeslib.as_array(D, shape=(nD,)).copy() if nD > 0 else np.zeros(0)
        img_pts, _ = cv2.projectPoints(pts, r, t, Km, Dv)
        np.ctypeslib.as_array(out, shape=(n, 2))[:] = img_pts.reshape(n, 2).astype(np.float32)   # std::vector<cv::Point2f>

    def draw(img, rows, cols, step, channels, what, geom, color, thickness):
        buf = np.ctypeslib.as_array(img, shape=(rows * step,))
        view = np.lib.stride_tricks.as_strided(buf, shape=(rows, cols, channels), strides=(step, channels, 1))
        col = tuple(float(color[i]) for i in range(4))
        g = [int(geom[i]) for i in range(4)]
        if what == 0:
            cv2.line(view, (g[0], g[1]), (g[2], g[3]), col, thickness)
        elif what == 1:
            cv2.circle(view, (g[0], g[1]), g[2], col, thickness)
        else:
            cv2.rectangle(view, (g[0], g[1], g[2], g[3]), col, thickness)

    cb = _Callbacks()
    fields = dict(_Callbacks._fields_)
    for name, fn in [("threshold", threshold), ("gaussian_blur", gaussian_blur), ("find_contours", find_contours),
                     ("contour_area", contour_area), ("bounding_rect", bounding_rect), ("moments", moments),
                     ("undistort_points", undistort_points), ("project_points", project_points), ("draw", draw)]:
        setattr(cb, name, fields[name](fn))
    return cb, state


def available() -> bool:
    """True when the library exists or can be built (needs /root/reference)."""
    if os.path.exists(_SO):
        return True
    if not os.path.isdir("/root/reference/monocular_pose_estimator_lib/src"):
        return False
    subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_pose.so is absent and cannot be built here")
        L = C.CDLL(_SO)
        L.mper_create.restype = C.c_void_p
        # same signatures as the mpeo_* functions (oracle/pose_oracle.py) ...
        for name, args, res in pose_oracle.SIGNATURES:
            rname = "mper_" + name[len("mpeo_"):]
            if hasattr(L, rname):
                fn = getattr(L, rname)
                fn.argtypes = args
                fn.restype = res
        # ... plus what only the reference build offers
        for name, args, res in [
            ("mper_set_cv_callbacks", [C.POINTER(_Callbacks)], None),
            ("mper_combinations_no_replacement", [C.c_uint, C.c_uint, up, C.c_int, ip], C.c_int),
            ("mper_permutations_no_replacement", [C.c_uint, C.c_uint, up, C.c_int, ip], C.c_int),
            ("mper_num_combinations", [C.c_uint, C.c_uint], C.c_uint),
            ("mper_num_permutations", [C.c_uint, C.c_uint], C.c_uint),
            ("mper_set_detector_params", [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint], None),
            ("mper_get_image_points", [C.c_void_p, dp], C.c_int),
            ("mper_correspondences_from_histogram", [C.c_void_p, up, C.c_int, C.c_int, up], C.c_int),
            ("mper_min_distances_and_pairs", [C.c_void_p, dp, C.c_int, dp, C.c_int, up, dp], None),
            ("mper_compute_jacobian", [dp, dp, dp, dp], None),
            ("mper_compute_transformation", [dp, dp, C.c_int, dp], None),
            ("mper_find_leds", [C.c_void_p, C.c_int, C.c_int, C.c_long, ip, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, dp, dp, C.c_int, dp, ip, fp], C.c_int),
            ("mper_estimate_body_pose", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_double], C.c_int),
            ("mper_get_roi", [C.c_void_p, ip], None),
            ("mper_get_distorted_centers", [C.c_void_p, fp], C.c_int),
            ("mper_create_visualization_image", [C.c_void_p, C.c_int, C.c_int, C.c_long, dp, dp, dp, C.c_int, ip, fp, C.c_int], None),
        ]:
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        cb, state = _make_callbacks()
        _KEEP.extend([cb, state])
        L.mper_set_cv_callbacks(C.byref(cb))
        _LIB = L
    return _LIB


class _AsOracle:
    """Presents libref_pose.so under the mpeo_* names so that PoseEstimatorOracle's accessors drive the reference build."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        if name.startswith("mpeo_"):
            return getattr(self._L, "mper_" + name[len("mpeo_"):])
        return getattr(self._L, name)


def _dpp(a):
    return a.ctypes.data_as(dp)


def combinations_no_replacement(N, K):
    out = np.zeros((4096, K), np.uint32)
    nc = C.c_int()
    n = lib().mper_combinations_no_replacement(N, K, out.ctypes.data_as(up), 4096, C.byref(nc))
    assert n >= 0
    return out.reshape(-1)[: n * nc.value].reshape(n, nc.value).copy()


def permutations_no_replacement(N, K):
    out = np.zeros(8192 * K, np.uint32)
    nc = C.c_int()
    n = lib().mper_permutations_no_replacement(N, K, out.ctypes.data_as(up), 8192, C.byref(nc))
    assert n >= 0
    return out[: n * nc.value].reshape(n, nc.value).copy()


def p3p(feature_vectors, world_points):
    f = np.ascontiguousarray(np.asarray(feature_vectors, np.float64).T)
    P = np.ascontiguousarray(np.asarray(world_points, np.float64).T)
    sol = np.zeros((4, 3, 4))
    rc = lib().mper_p3p(_dpp(f), _dpp(P), _dpp(sol))
    return rc, sol


def exponential_map(twist):
    t = np.ascontiguousarray(twist, np.float64); out = np.zeros((4, 4))
    lib().mper_exponential_map(_dpp(t), _dpp(out)); return out


def logarithm_map(T):
    T = np.ascontiguousarray(T, np.float64); out = np.zeros(6)
    lib().mper_logarithm_map(_dpp(T), _dpp(out)); return out


def compute_jacobian(T, point, focal):
    T = np.ascontiguousarray(T, np.float64); p = np.ascontiguousarray(point, np.float64)
    f = np.ascontiguousarray(focal, np.float64); out = np.zeros((2, 6))
    lib().mper_compute_jacobian(_dpp(T), _dpp(p), _dpp(f), _dpp(out)); return out


def compute_transformation(object_pts, reprojected_pts):
    a = np.ascontiguousarray(object_pts, np.float64); b = np.ascontiguousarray(reprojected_pts, np.float64)
    out = np.zeros((4, 4))
    lib().mper_compute_transformation(_dpp(a), _dpp(b), len(a), _dpp(out)); return out


def find_leds(image, roi, threshold_value, gaussian_sigma, min_blob_area, max_blob_area, max_width_height_distortion,
              max_circular_distortion, K, D, pixel_positions=None):
    """LEDDetector::findLeds of the reference build.  Returns (pixel_positions, distorted_centres) with the reference's
    'untouched when nothing was found' behaviour for pixel_positions (None in, None out)."""
    img = np.ascontiguousarray(image, np.uint8)
    K = np.ascontiguousarray(K, np.float64); D = np.ascontiguousarray(D, np.float64)
    px = np.zeros((256, 2)); n_px = C.c_int(0)
    if pixel_positions is not None:
        n_px.value = len(pixel_positions); px[: len(pixel_positions)] = pixel_positions
    centers = np.zeros((256, 2), np.float32)
    r = (C.c_int * 4)(*[int(v) for v in roi])
    n = lib().mper_find_leds(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], img.strides[0], r,
                             int(threshold_value), gaussian_sigma, min_blob_area, max_blob_area,
                             max_width_height_distortion, max_circular_distortion, _dpp(K), _dpp(D), len(D), _dpp(px),
                             C.byref(n_px), centers.ctypes.data_as(fp))
    out_px = px[: n_px.value].copy() if (n > 0 or pixel_positions is not None) else None
    return out_px, centers[:n].copy()


def create_visualization_image(image_bgr, pose, K, D, roi, centers):
    """Visualization::createVisualizationImage of the reference build (visualization.cpp:57-104) on a copy of a 3-channel image:
    its projectPoints / line / circle / rectangle calls reach the cv2 functions of the same name."""
    img = np.ascontiguousarray(image_bgr, np.uint8).copy()
    assert img.ndim == 3 and img.shape[2] == 3
    pose = np.ascontiguousarray(pose, np.float64).reshape(16); K = np.ascontiguousarray(K, np.float64).reshape(9)
    D = np.ascontiguousarray(D, np.float64)
    c = np.ascontiguousarray(centers, np.float32).reshape(-1, 2)
    r = (C.c_int * 4)(*[int(v) for v in roi])
    lib().mper_create_visualization_image(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], img.strides[0], _dpp(pose), _dpp(K),
                                          _dpp(D), len(D), r, c.ctypes.data_as(fp), len(c))
    return img


class PoseEstimatorRef(pose_oracle.PoseEstimatorOracle):
    """The reference's PoseEstimator (unmodified source) behind the accessors of PoseEstimatorOracle.
    estimate_body_pose runs the reference's own estimateBodyPose (pose_estimator.cpp:62-147), not the Python restatement."""

    def __init__(self, K, D, markers, params):
        super().__init__(K, D, markers, params, _lib=_AsOracle(lib()))
        p = params
        self.L.mper_set_detector_params(self.h, int(p.threshold_value), p.gaussian_sigma, p.min_blob_area, p.max_blob_area,
                                        p.max_width_height_distortion, p.max_circular_distortion,
                                        int(p.roi_border_thickness))

    def counters(self):
        raise NotImplementedError("the unmodified reference has no counters")

    def estimate_body_pose(self, image, time_to_predict):
        img = np.ascontiguousarray(image, np.uint8)
        ok = self.L.mper_estimate_body_pose(self.h, img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1],
                                            img.strides[0], float(time_to_predict))
        roi = (C.c_int * 4)(); self.L.mper_get_roi(self.h, roi)
        self.region_of_interest = tuple(roi)
        c = np.zeros((256, 2), np.float32)
        n = self.L.mper_get_distorted_centers(self.h, c.ctypes.data_as(fp))
        self.distorted_detection_centers = c[:n].copy()
        pts = np.zeros((256, 2))
        self.n_det = self.L.mper_get_image_points(self.h, _dpp(pts))
        self.pose_updated = bool(ok)
        return self.pose_updated
