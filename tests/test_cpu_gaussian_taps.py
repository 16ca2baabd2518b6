"""The fixed-point Gaussian taps the library derives from gaussian_sigma (csrc/mpe_abi.cu gaussian_taps_8u, exported host-only as
mpe_debug_gaussian_taps) against OpenCV itself, for a dense sweep of sigma over the whole dynamic_reconfigure range (0, 6]
(monocular_pose_estimator/cfg/MonocularPoseEstimator.cfg:13): the 8.8 -> 16.16 fixed-point blur evaluated with the library's taps must
reproduce cv2.GaussianBlur(ksize=(0,0)) byte for byte, so a rounding tie at some sigma that flipped a tap would show here.
No GPU needed: the call computes on the host."""
import ctypes as C

import cv2
import numpy as np

from rpg_monocular_pose_estimator_b200 import _lib


def lib_taps(L, sigma):
    r = C.c_int(0)
    t = (C.c_uint32 * 64)()
    rc = L.mpe_debug_gaussian_taps(float(sigma), C.byref(r), t, 64)
    return rc, r.value, np.array(t[: 2 * r.value + 1], np.int64)


def refl(i, N):
    if N == 1:
        return 0
    while i < 0 or i >= N:
        i = -i if i < 0 else 2 * (N - 1) - i
    return i


def model_blur(img, t):
    n = len(t); r = n // 2
    H, W = img.shape
    xi = np.array([[refl(x + k - r, W) for k in range(n)] for x in range(W)])
    yi = np.array([[refl(y + k - r, H) for k in range(n)] for y in range(H)])
    h = (img.astype(np.int64)[:, xi] * t).sum(-1)
    v = (h[yi, :] * t[None, :, None]).sum(1)
    return ((v + 32768) >> 16).astype(np.uint8)


def test_taps_reproduce_cv2_for_a_dense_sigma_sweep():
    L = _lib.load_library()
    rng = np.random.default_rng(5)
    sigmas = np.concatenate([np.linspace(0.01, 6.0, 660), rng.uniform(0.01, 6.0, 120), [0.3, 0.6, 1.0, 1.5, 2.0, 3.0, 6.0]])
    img = cv2.threshold(rng.integers(0, 256, (48, 64), dtype=np.uint8), 140, 255, cv2.THRESH_TOZERO)[1]
    impulse = np.zeros((81, 81), np.uint8); impulse[40, 40] = 255
    n_checked = 0
    for sigma in sigmas:
        rc, r, t = lib_taps(L, sigma)
        ksize = max(int(np.rint(sigma * 6 + 1)) | 1, 3)      # a 1-tap kernel (identity) is carried as 0 256 0
        assert rc == 0 and 2 * r + 1 == ksize, (sigma, rc, r, ksize)
        assert t.sum() == 256 and np.array_equal(t, t[::-1])
        assert np.array_equal(model_blur(img, t), cv2.GaussianBlur(img.copy(), (0, 0), float(sigma), sigmaY=float(sigma))), sigma
        assert np.array_equal(model_blur(impulse, t), cv2.GaussianBlur(impulse.copy(), (0, 0), float(sigma), sigmaY=float(sigma))), sigma
        n_checked += 1
    assert n_checked >= 700


def test_out_of_range_sigma_is_rejected():
    L = _lib.load_library()
    for sigma in (0.0, -1.0, float("nan"), 6.3, 50.0):
        rc, _, _ = lib_taps(L, sigma)
        assert rc != 0, sigma
