"""K2 tier 1 (csrc/p3p_tier1.cuh) on the CPU: the host build of the pre-test against the host build of the product's exact
P3P path (oracle/tier1_check.cpp), over every P3P problem of many frames.  What must hold for the two-tier sweep to leave
the histogram unchanged: no problem that votes under the exact arithmetic is rejected (violations == 0), rejected problems
keep a distance to the tolerance (closest call), and on problems tier 1 does not flag its roots and back-projections agree
with the exact ones to a tiny fraction of the margin."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
MARGIN = 0.25      # pixels; mpe_abi.cu run_sweep


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("problems", "conditioned", "survivors", "flagged", "voting_problems", "violations")] + \
               [("closest_call", C.c_double), ("max_root_dev", C.c_double), ("max_pixel_dev", C.c_double), ("compared", C.c_longlong)]


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ORACLE, "libtier1_check.so")
    if not os.path.isdir("/usr/local/cuda/include") and not os.path.exists(so):
        pytest.skip("no CUDA headers to build the host copy of the device code")
    subprocess.check_call(["make", "-C", ORACLE, "libtier1_check.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.t1c_frame.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    L.t1c_init.argtypes = [C.POINTER(Stats)]
    L.t1c_set_fp32.argtypes = [C.c_int]
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run(L, K, mk, dets, tol, margin=MARGIN):
    S = Stats(); L.t1c_init(C.byref(S))
    K = np.ascontiguousarray(K); mk = np.ascontiguousarray(mk, np.float64)
    for det in dets:
        det = np.ascontiguousarray(det, np.float64)
        L.t1c_frame(_dp(K), _dp(mk), len(mk), _dp(det), len(det), tol, margin, C.byref(S))
    return S


def views(rng, K, D, mk, n, noise=0.3, junk=0):
    out = []
    for _ in range(n):
        T = synth.sample_pose(rng, K, D, mk, 752, 480, z_range=(0.8, 1.0) if len(mk) == 8 else (0.4, 1.2))
        _, und, _ = synth.project_distorted(K, D, T, mk)
        det = und + rng.normal(size=und.shape) * noise
        if junk:
            det = np.vstack([det, np.stack([rng.uniform(0, 752, junk), rng.uniform(0, 480, junk)], 1)])
        out.append(det[rng.permutation(len(det))])
    return out


# fp32 = the optional build (MPE_T1_FP32=1: back-projection test in single precision), fp64 = the default build, everything in double
DEV_LIMIT = {1: (MARGIN * 4e-2, MARGIN * 0.2), 0: (MARGIN * 1e-2, MARGIN * 4e-2)}      # px: object views, adversarial inputs


@pytest.mark.parametrize("fp32", [1, 0])
@pytest.mark.parametrize("n_leds,n_frames", [(4, 1500), (5, 800), (8, 12)])
def test_tier1_is_conservative_on_object_views(lib, n_leds, n_frames, fp32):
    lib.t1c_set_fp32(fp32)
    rng = np.random.default_rng(n_leds)
    K, D = synth.camera()
    mk = synth.markers(n_leds)
    S = run(lib, K, mk, views(rng, K, D, mk, n_frames), tol=5.0)
    print(f"fp32={fp32} n_leds={n_leds}: {S.problems} problems, {S.conditioned} conditioned, survivors {S.survivors / S.problems:.3%} "
          f"(flagged {S.flagged / S.problems:.3%}), voting {S.voting_problems / S.problems:.3%}, closest call {S.closest_call:.3f} px, "
          f"root dev {S.max_root_dev:.2e}, pixel dev {S.max_pixel_dev:.2e} px over {S.compared} hypotheses")
    assert S.violations == 0
    assert S.closest_call >= MARGIN * 0.99          # a rejected problem is at least ~margin away from voting
    assert S.max_root_dev < 1e-7 and S.max_pixel_dev < DEV_LIMIT[fp32][0]     # measured: ~1e-8; 3e-3 px (fp32) / 4e-4 px (fp64)
    assert S.survivors < 0.25 * S.problems and S.voting_problems > 0


@pytest.mark.parametrize("fp32", [1, 0])
def test_tier1_is_conservative_on_adversarial_inputs(lib, fp32):
    lib.t1c_set_fp32(fp32)
    rng = np.random.default_rng(99)
    K, D = synth.camera()
    total = 0
    for case in range(12):
        n = int(rng.choice([4, 5, 6]))
        mk = rng.uniform(-0.15, 0.15, (n, 3))
        if case % 4 == 1:
            mk[2] = mk[0] + 0.6 * (mk[1] - mk[0]) + rng.normal(size=3) * 1e-9      # nearly colinear triple
        if case % 4 == 2:
            mk[:, 2] = 0.0                                                          # planar object
        dets = views(rng, K, D, mk, 6, noise=0.5, junk=int(rng.integers(0, 3)))
        dets += [np.stack([rng.uniform(0, 752, n + 1), rng.uniform(0, 480, n + 1)], 1) for _ in range(6)]     # junk only
        d = dets[0].copy(); d = np.vstack([d, d[0] + [1.0, 0.0]]); dets.append(d)                              # detections 1 px apart
        for tol in (0.01, 1.0, 5.0, 60.0):
            S = run(lib, K, mk, dets, tol=tol)
            assert S.violations == 0, (case, tol)
            assert S.closest_call >= MARGIN * 0.99, (case, tol, S.closest_call)
            assert S.max_root_dev < 1e-6 and S.max_pixel_dev < DEV_LIMIT[fp32][1], (case, tol, S.max_root_dev, S.max_pixel_dev)
            total += S.problems
    assert total > 100000


@pytest.mark.parametrize("fp32", [1, 0])
@pytest.mark.parametrize("name,focal_scale,size,z_range,object_scale", [
    ("telephoto", 8.0, (752, 480), (4.0, 9.0), 1.0),            # huge fx: pixel errors scale with the focal length
    ("wide", 1.0 / 3.0, (752, 480), (0.15, 0.4), 1.0),          # close range, strong perspective
    ("1080p", None, (1920, 1080), (0.4, 1.2), 1.0),             # BASELINE config 3's camera (with its distortion)
    ("far", 1.0, (752, 480), (3.0, 6.0), 1.0),                  # LEDs a few pixels apart: many hypotheses within the tolerance
    ("large_object", 1.0, (752, 480), (3.0, 6.0), 8.0),
])
def test_tier1_is_conservative_for_other_cameras_and_object_sizes(lib, fp32, name, focal_scale, size, z_range, object_scale):
    """The margin (0.25 px) and the flags were calibrated on the bench camera; the pre-test must stay conservative when the focal
    length, the image size, the range or the object size change by an order of magnitude."""
    lib.t1c_set_fp32(fp32)
    W, H = size
    if focal_scale is None:
        K, D = synth.camera(W, H)
    else:
        K, D = synth.camera()
        K = K.copy(); K[0, 0] *= focal_scale; K[1, 1] *= focal_scale
        D = np.zeros(5)
    mk = synth.markers(5) * object_scale
    rng = np.random.default_rng(5)
    dets = []
    for _ in range(100):
        T = synth.sample_pose(rng, K, D, mk, W, H, z_range=z_range)
        _, und, _ = synth.project_distorted(K, D, T, mk)
        det = und + rng.normal(size=und.shape) * 0.3
        dets.append(det[rng.permutation(len(det))])
    for tol in (1.0, 5.0):
        S = run(lib, K, mk, dets, tol=tol)
        assert S.violations == 0, (name, tol)
        assert S.closest_call >= MARGIN * 0.99, (name, tol, S.closest_call)
        assert S.max_root_dev < 1e-7 and S.max_pixel_dev < DEV_LIMIT[fp32][1], (name, tol, S.max_root_dev, S.max_pixel_dev)
        assert S.voting_problems > 0 and S.survivors < 0.3 * S.problems
