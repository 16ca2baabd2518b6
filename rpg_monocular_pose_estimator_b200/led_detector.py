"""Host-side mirror of monocular_pose_estimator::LEDDetector
(/root/reference/monocular_pose_estimator_lib/include/monocular_pose_estimator_lib/led_detector.h:45-140)."""
from __future__ import annotations

import numpy as np


class LEDDetector:
    @staticmethod
    def findLeds(image, ROI, threshold_value=None, gaussian_sigma=None, min_blob_area=None, max_blob_area=None,
                 max_width_height_distortion=None, max_circular_distortion=None, camera_matrix_K=None,
                 camera_distortion_coeffs=None, *, context):
        """led_detector.h:84-88 / led_detector.cpp:35-112 on the GPU (K1).  Returns (pixel_positions n x 2 float64,
        distorted_detection_centers n x 2 float32, flags).  If the tunables / camera are given they are pushed to the
        context first; otherwise the context's configuration is used (PoseEstimator does that)."""
        if threshold_value is not None:
            from .synth import Params
            p = Params(int(threshold_value), float(gaussian_sigma), float(min_blob_area), float(max_blob_area),
                       float(max_width_height_distortion), float(max_circular_distortion))
            context.set_params(p)
        if camera_matrix_K is not None:
            context.set_camera(camera_matrix_K, camera_distortion_coeffs)
        return context.find_leds(image, ROI)

    @staticmethod
    def determineROI(pixel_positions, image_size, border_size, camera_matrix_K, camera_distortion_coeffs):
        """led_detector.h:105-106 / led_detector.cpp:114-179 (with distortPoints :181-224: float32 corners, double arithmetic).
        image_size = (width, height).  Returns (x, y, w, h).  Computed by the library's host helper — the function the device
        loop runs per stream (csrc/tracking_math.cuh)."""
        import ctypes as C
        from . import _lib
        L = _lib.load_library()
        px = np.ascontiguousarray(np.asarray(pixel_positions, np.float64).reshape(-1, 2))
        K = np.ascontiguousarray(camera_matrix_K, np.float64)
        D = np.ascontiguousarray(camera_distortion_coeffs, np.float64)
        dp = C.POINTER(C.c_double)
        r = _lib.MpeRect()
        rc = L.mpe_host_determine_roi(px.ctypes.data_as(dp), len(px), int(image_size[0]), int(image_size[1]), int(border_size),
                                      K.ctypes.data_as(dp), D.ctypes.data_as(dp), len(D), C.byref(r))
        if rc != 0:
            raise _lib.MpeError(f"mpe_host_determine_roi failed ({rc})")
        return (r.x, r.y, r.width, r.height)
