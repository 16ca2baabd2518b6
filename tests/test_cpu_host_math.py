"""The library's host-side tracking helpers (mpe_host_*: the functions of csrc/tracking_math.cuh that K4 runs per stream on the GPU,
compiled for the host) against the CPU oracle — bit for bit, since both run on glibc — and, where the reference build exists
(oracle/_ref), against the unmodified reference sources.  These are what the stage-by-stage mirrors (C++ shim, Python class) call,
so this pins predictPose / predictMarkerPositionsInImage / determineROI of stage mode to the reference's operation order.
No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import _lib, synth
from rpg_monocular_pose_estimator_b200.led_detector import LEDDetector
from oracle import pose_oracle, ref_pose
from tests.test_oracle_pose_ref import same_bits

dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(dp)


def estimators():
    K, D = synth.camera()
    m = synth.markers(5)
    out = [pose_oracle.PoseEstimatorOracle(K, D, m, synth.Params())]
    if ref_pose.available():
        out.append(ref_pose.PoseEstimatorRef(K, D, m, synth.Params()))
    return K, D, m, out


def random_pose(rng, scale=1.0):
    T = np.eye(4)
    T[:3, :3] = synth.rodrigues(rng.normal(size=3) * 0.4 * scale)
    T[:3, 3] = [rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(0.4, 1.5)]
    return T


def test_exp_log_maps_bit_identical():
    L = _lib.load_library()
    rng = np.random.default_rng(1)
    for i in range(1500):
        tw = rng.normal(size=6) * rng.choice([1e-12, 1e-6, 1e-2, 1.0, 3.0])
        if i % 10 == 0:
            tw[3:] = 0
        a = np.zeros((4, 4)); b = np.zeros((4, 4))
        assert L.mpe_host_exponential_map(_p(np.ascontiguousarray(tw)), _p(a)) == 0
        pose_oracle.lib().mpeo_exponential_map(_p(np.ascontiguousarray(tw)), _p(b))
        assert same_bits(a, b), i
        T = a.copy()
        if i % 13 == 0:
            T[:3, 3] = 0
        x = np.zeros(6); y = np.zeros(6)
        assert L.mpe_host_logarithm_map(_p(np.ascontiguousarray(T)), _p(x)) == 0
        pose_oracle.lib().mpeo_logarithm_map(_p(np.ascontiguousarray(T)), _p(y))
        assert same_bits(x, y), i


def test_predict_pose_project_markers_determine_roi_bit_identical():
    L = _lib.load_library()
    rng = np.random.default_rng(2)
    K, D, m, ests = estimators()
    for i in range(400):
        prev = random_pose(rng)
        d = np.eye(4); d[:3, :3] = synth.rodrigues(rng.normal(size=3) * 0.02); d[:3, 3] = rng.normal(size=3) * 0.005
        cur = prev @ d
        t_prev, t_cur, t_pred = 1.0, 1.0 + 1 / 60.0, 1.0 + 2 / 60.0 + rng.uniform(0, 0.01)
        out = np.zeros((4, 4))
        assert L.mpe_host_predict_pose(_p(np.ascontiguousarray(prev)), _p(np.ascontiguousarray(cur)), t_prev, t_cur, t_pred, _p(out)) == 0
        px = np.zeros((5, 2))
        assert L.mpe_host_project_markers(_p(np.ascontiguousarray(K)), _p(out), _p(np.ascontiguousarray(m)), 5, _p(px)) == 0
        border = int(rng.integers(0, 40))
        roi = LEDDetector.determineROI(px, (752, 480), border, K, D)
        for e in ests:
            e.L.mpeo_set_state(e.h, _p(np.ascontiguousarray(cur)), _p(np.ascontiguousarray(prev)), t_cur, t_prev, 2)
            e.L.mpeo_predict_pose(e.h, t_pred)
            e.L.mpeo_predict_marker_positions(e.h)
            is_ref = not isinstance(e.L, C.CDLL)             # the reference build inverts previous_pose_ with the stand-in's cofactor formula
            if is_ref:
                assert np.abs(e.predicted_pose() - out).max() < 1e-13, i
                assert np.abs(e.predicted_pixels() - px).max() < 1e-9, i
            else:
                assert same_bits(e.predicted_pose(), out), i
                assert same_bits(e.predicted_pixels(), px), i
            assert e.determine_roi(752, 480) == roi or border != e.params.roi_border_thickness
            r = (C.c_int * 4)()
            e.L.mpeo_set_predicted_pixels(e.h, _p(np.ascontiguousarray(px)), 5)
            e.L.mpeo_determine_roi(e.h, 752, 480, border, r)
            assert tuple(r) == roi, (i, tuple(r), roi)
    # degenerate predictions far to the right of the image: the clamped extent is < 1 -> whole image (led_detector.cpp:166-169)
    assert LEDDetector.determineROI(np.full((5, 2), 5000.0), (752, 480), 20, K, D) == (0, 0, 752, 480)


def test_pose_estimator_augment_image_draws_the_last_result():
    """PoseEstimator.augmentImage (pose_estimator.cpp:44-48) only touches host state: checked here without a device by handing the
    mirror a placeholder context."""
    from rpg_monocular_pose_estimator_b200.pose_estimator import PoseEstimator
    from rpg_monocular_pose_estimator_b200.visualization import Visualization
    from rpg_monocular_pose_estimator_b200 import synth
    K, D = synth.camera(752, 480)
    est = PoseEstimator(context=object())
    est.camera_matrix_K_, est.camera_distortion_coeffs_ = K, D
    est.predicted_pose_ = np.eye(4); est.predicted_pose_[:3, 3] = [0.03, -0.02, 0.7]
    est.region_of_interest_ = (200, 150, 180, 120)
    est.distorted_detection_centers_ = np.array([[250.5, 200.25], [300.0, 210.0]], np.float32)
    img = np.zeros((480, 752, 3), np.uint8)
    out = est.augmentImage(img)
    assert out is img and img.any()
    want = Visualization.createVisualizationImage(np.zeros((480, 752, 3), np.uint8), est.predicted_pose_, K, D, est.region_of_interest_,
                                                  est.distorted_detection_centers_)
    assert np.array_equal(img, want)
    assert (img[150, 200:380] == (255, 0, 0)).all()            # the ROI rectangle is blue in BGR
