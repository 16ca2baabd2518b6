"""Host-side mirror of monocular_pose_estimator::LEDDetector
(/root/reference/monocular_pose_estimator_lib/include/monocular_pose_estimator_lib/led_detector.h:45-140)."""
from __future__ import annotations

import math
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
    def distortPoints(src, camera_matrix_K, distortion_matrix):
        """led_detector.cpp:181-224 (float32 in, float32 out, double arithmetic; reads D[4] unconditionally)."""
        K = np.asarray(camera_matrix_K, np.float64)
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        k1, k2, p1, p2, k3 = [float(v) for v in distortion_matrix[:5]]
        out = []
        for (px, py) in src:
            x = (float(np.float32(px)) - cx) / fx
            y = (float(np.float32(py)) - cy) / fy
            r2 = x * x + y * y
            xc = x * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2)
            yc = y * (1. + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2)
            xc = xc + (2. * p1 * x * y + p2 * (r2 + 2. * x * x))
            yc = yc + (p1 * (r2 + 2. * y * y) + 2. * p2 * x * y)
            out.append((np.float32(xc * fx + cx), np.float32(yc * fy + cy)))
        return out

    @staticmethod
    def determineROI(pixel_positions, image_size, border_size, camera_matrix_K, camera_distortion_coeffs):
        """led_detector.h:105-106 / led_detector.cpp:114-179.  image_size = (width, height).  Returns (x, y, w, h)."""
        x_min, x_max, y_min, y_max = math.inf, 0.0, math.inf, 0.0
        for p in pixel_positions:
            if p[0] < x_min: x_min = float(p[0])
            if p[0] > x_max: x_max = float(p[0])
            if p[1] < y_min: y_min = float(p[1])
            if p[1] > y_max: y_max = float(p[1])
        with np.errstate(over="ignore"):
            corners = [(np.float32(x_min), np.float32(y_min)), (np.float32(x_max), np.float32(y_max))]
        d = LEDDetector.distortPoints(corners, camera_matrix_K, camera_distortion_coeffs)
        x_min_d, y_min_d, x_max_d, y_max_d = float(d[0][0]), float(d[0][1]), float(d[1][0]), float(d[1][1])
        W, H = image_size
        x0 = max(0.0, min(float(W), x_min_d - border_size)); x1 = max(0.0, min(float(W), x_max_d + border_size))
        y0 = max(0.0, min(float(H), y_min_d - border_size)); y1 = max(0.0, min(float(H), y_max_d + border_size))
        if x1 - x0 < 1 or y1 - y0 < 1 or math.isnan(x1 - x0) or math.isnan(y1 - y0):
            return (0, 0, int(W), int(H))
        return (int(x0), int(y0), int(x1 - x0), int(y1 - y0))
