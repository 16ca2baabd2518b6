"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import cv2

from rpg_monocular_pose_estimator_b200 import synth
from oracle import find_leds_cv2, pose_oracle


def oracle_find_leds(image, roi, params, K, D, debug=False):
    return find_leds_cv2.find_leds(image, roi, params.threshold_value, params.gaussian_sigma, params.min_blob_area,
                                   params.max_blob_area, params.max_width_height_distortion, params.max_circular_distortion,
                                   K, D, return_debug=debug)


def random_blob_image(rng, h, w, n_blobs=8, kind="ellipse", noise_max=20):
    """Ellipses / rings / dots of varied size and brightness over dark noise: exercises the filters, merged blobs,
    border-touching blobs and nested components."""
    img = rng.integers(0, noise_max, size=(h, w), dtype=np.uint8)
    for _ in range(n_blobs):
        cx, cy = int(rng.integers(-4, w + 4)), int(rng.integers(-4, h + 4))
        ax, ay = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        val = int(rng.integers(120, 256))
        k = rng.integers(0, 10)
        if kind == "mixed" and k < 2:
            cv2.ellipse(img, (cx, cy), (ax + 6, ay + 6), float(rng.uniform(0, 180)), 0, 360, val, int(rng.integers(1, 4)))   # ring
            if rng.random() < 0.7:
                cv2.circle(img, (cx, cy), int(rng.integers(1, 3)), val, -1)                                                   # dot inside
        elif kind == "mixed" and k < 3:
            img[max(cy, 0):cy + 2, max(cx, 0):cx + 2] = val                                                                 # tiny speck
        else:
            cv2.ellipse(img, (cx, cy), (ax, ay), float(rng.uniform(0, 180)), 0, 360, val, -1)
    return img


def pose_error(Ta, Tb):
    """translation error (m), rotation error (rad)"""
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    R = Ta[:3, :3].T @ Tb[:3, :3]
    if not np.all(np.isfinite(R)):
        return float("nan"), float("nan")
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = float(np.linalg.norm(w))                       # sin(angle): accurate for small angles (arccos of the trace is not)
    c = (np.trace(R) - 1) / 2
    return dt, float(np.arctan2(s, c))
