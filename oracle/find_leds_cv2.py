"""TEST INFRASTRUCTURE — findLeds oracle.

Line-by-line restatement of LEDDetector::findLeds
(/root/reference/monocular_pose_estimator_lib/src/led_detector.cpp:35-112) on top of the SAME OpenCV
kernels the reference calls, through the official `cv2` bindings (opencv-python-headless 4.13.0 in this
image).  The arithmetic of this stage lives in OpenCV (third-party, version not pinned by the reference:
L/CMakeLists.txt:10 `find_package(OpenCV REQUIRED)`), so cv2 4.13.0 IS the contract here.

Parity status: pinned to cv2 4.13.0 run in this container (tests/golden/find_leds_*.npz are its outputs;
generator: tests/golden/make_golden.py).  The reference itself ships no golden vectors.
"""
from __future__ import annotations

import math
import numpy as np
import cv2


def find_leds(image: np.ndarray, roi, threshold_value: int, gaussian_sigma: float, min_blob_area: float,
              max_blob_area: float, max_width_height_distortion: float, max_circular_distortion: float,
              K: np.ndarray, D: np.ndarray, return_debug: bool = False):
    """Returns (pixel_positions float64 n x 2 or None when nothing was found [reference leaves the output
    untouched, led_detector.cpp:91], distorted_detection_centers float32 n x 2)."""
    x, y, w, h = [int(v) for v in roi]
    sub = image[y:y + h, x:x + w]
    # :44  cv::threshold(image(ROI), bw_image, threshold_value, 255, cv::THRESH_TOZERO)
    _, bw = cv2.threshold(sub, threshold_value, 255, cv2.THRESH_TOZERO)
    # :48-51  GaussianBlur(bw_image.clone(), gaussian_image, Size(0,0), sigma, sigma, BORDER_DEFAULT)
    gaussian = cv2.GaussianBlur(np.ascontiguousarray(bw).copy(), (0, 0), gaussian_sigma, sigmaY=gaussian_sigma,
                                borderType=cv2.BORDER_DEFAULT)
    # :57  findContours(gaussian_image.clone(), contours, CV_RETR_EXTERNAL, CV_CHAIN_APPROX_NONE)
    contours, _ = cv2.findContours(gaussian.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)

    distorted_points = []
    debug = []
    for c in contours:
        area = cv2.contourArea(c)                        # :67
        rx, ry, rw, rh = cv2.boundingRect(c)             # :68
        mu = cv2.moments(c, False)                       # :71-72
        with np.errstate(divide="ignore", invalid="ignore"):
            mcx = np.float32(np.float64(mu["m10"]) / np.float64(mu["m00"]))   # Point2f(mu.m10/mu.m00, ...)
            mcy = np.float32(np.float64(mu["m01"]) / np.float64(mu["m00"]))
        mc = (np.float32(mcx + np.float32(x)), np.float32(mcy + np.float32(y)))  # + Point2f(ROI.x, ROI.y) :74
        with np.errstate(divide="ignore", invalid="ignore"):
            wh = abs(1 - min(np.float64(rw) / np.float64(rh), np.float64(rh) / np.float64(rw)))
            cw = abs(1 - (np.float64(area) / (math.pi * np.float64((rw // 2) ** 2))))   # rect.width / 2 is integer division :80
            ch = abs(1 - (np.float64(area) / (math.pi * np.float64((rh // 2) ** 2))))
        keep = (area >= min_blob_area and area <= max_blob_area and wh <= max_width_height_distortion
                and cw <= max_circular_distortion and ch <= max_circular_distortion)   # :77-81
        debug.append(dict(area=area, rect=(rx, ry, rw, rh), mc=mc, keep=bool(keep), start=tuple(c[0][0])))
        if keep:
            distorted_points.append(mc)

    centers = np.array(distorted_points, dtype=np.float32).reshape(-1, 2)          # :89
    pixel_positions = None
    if len(distorted_points) > 0:                                                   # :91
        und = cv2.undistortPoints(centers.reshape(-1, 1, 2), K, D, None, K)          # :97-98
        pixel_positions = und.reshape(-1, 2).astype(np.float64)                     # :101-110
    if return_debug:
        return pixel_positions, centers, dict(blurred=gaussian, contours=debug)
    return pixel_positions, centers
