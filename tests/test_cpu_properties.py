"""Size-independent properties of the pose path, checked on the CPU oracle (no GPU): they hold for the reference's algorithm by
construction and would expose a mis-stated formula that the golden vectors (which the oracle itself wrote) cannot."""
import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from oracle import pose_oracle
from tests.helpers import pose_error


def _random_pose(rng, z=(0.5, 1.2)):
    T = np.eye(4)
    T[:3, :3] = synth.rodrigues(rng.normal(size=3) * 0.6)
    T[:3, 3] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.15, 0.15), rng.uniform(*z)]
    return T


def _project(K, T, pts):
    cam = (T[:3, :3] @ pts.T).T + T[:3, 3]
    return np.stack([K[0, 0] * cam[:, 0] / cam[:, 2] + K[0, 2], K[1, 1] * cam[:, 1] / cam[:, 2] + K[1, 2]], axis=1), cam


def test_p3p_is_equivariant_under_rigid_motions_of_the_world_frame():
    """P3P::computePoses (p3p.cpp:65-236) returns camera->world poses [R|C]: moving the world points by (R0, t0) must move every
    solution to [R0 R | R0 C + t0]; the true camera pose is among the four."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(200):
        pts = rng.uniform(-0.2, 0.2, size=(3, 3))                      # rows = world points
        T = _random_pose(rng)                                          # world -> camera
        cam = (T[:3, :3] @ pts.T).T + T[:3, 3]
        f = cam / np.linalg.norm(cam, axis=1, keepdims=True)
        rc, sol = pose_oracle.p3p(f.T, pts.T)
        assert rc == 0
        # the true camera->world pose is one of the solutions
        Rt, Ct = T[:3, :3].T, -T[:3, :3].T @ T[:3, 3]
        err = [np.abs(s[:, :3] - Rt).max() + np.abs(s[:, 3] - Ct).max() for s in sol if np.isfinite(s).all()]
        assert min(err) < 1e-7, min(err)
        R0 = synth.rodrigues(rng.normal(size=3)); t0 = rng.uniform(-1, 1, size=3)
        pts2 = (R0 @ pts.T).T + t0
        rc2, sol2 = pose_oracle.p3p(f.T, pts2.T)
        assert rc2 == 0
        for s, s2 in zip(sol, sol2):
            if not (np.isfinite(s).all() and np.isfinite(s2).all()):
                continue
            worst = max(worst, np.abs(R0 @ s[:, :3] - s2[:, :3]).max(), np.abs(R0 @ s[:, 3] + t0 - s2[:, 3]).max())
    assert worst < 1e-6, worst


@pytest.mark.parametrize("n_leds", [4, 5, 8])
def test_exact_detections_recover_the_pose_and_the_identity_permutation(n_leds):
    """Noise-free pinhole projections of the markers in marker order: initialise() must decode LED i <-> detection i for every LED
    and optimisePose must land on the pose that generated them (residual 0), from the Kabsch start, in a handful of iterations."""
    rng = np.random.default_rng(100 + n_leds)
    K, D = synth.camera()
    markers = synth.make_cold_scene(1, n_leds=n_leds, seed=1).markers
    params = synth.Params()
    for _ in range(6 if n_leds < 8 else 2):
        T = _random_pose(rng, z=(0.7, 1.0))
        px, cam = _project(K, T, markers)
        if cam[:, 2].min() < 0.3 or px.min() < 20 or px[:, 0].max() > 730 or px[:, 1].max() > 460:
            continue
        est = pose_oracle.PoseEstimatorOracle(K, D, markers, params)
        est.set_image_points(px)
        assert est.initialise() == 1
        corr = est.correspondences()
        assert len(corr) == n_leds and np.array_equal(corr[np.argsort(corr[:, 0])], np.stack([np.arange(1, n_leds + 1)] * 2, axis=1))
        it = est.optimise_pose()
        dt, dr = pose_error(est.predicted_pose(), T)
        assert dt < 1e-8 and dr < 1e-8, (dt, dr)
        assert 1 <= it <= 12


def test_gauss_newton_converges_to_the_same_pose_from_perturbed_starts():
    """optimisePose (pose_estimator.cpp:733-792) is a fixed-point iteration on the reprojection error: small perturbations of the
    start pose end in the same pose (to 1e-9) and the covariance is symmetric positive definite."""
    rng = np.random.default_rng(9)
    K, D = synth.camera()
    markers = synth.make_cold_scene(1, n_leds=5, seed=1).markers
    params = synth.Params()
    T = _random_pose(rng, z=(0.7, 1.0))
    px, _ = _project(K, T, markers)
    px = px + rng.normal(size=px.shape) * 0.2
    corr = np.stack([np.arange(1, 6)] * 2, axis=1).astype(np.uint32)
    finals = []
    for k in range(8):
        est = pose_oracle.PoseEstimatorOracle(K, D, markers, params)
        est.set_image_points(px)
        est.set_correspondences(corr)
        T0 = T.copy()
        if k:
            T0[:3, :3] = synth.rodrigues(rng.normal(size=3) * 0.03) @ T0[:3, :3]
            T0[:3, 3] += rng.normal(size=3) * 0.01
        est.set_predicted_pose(T0)
        it = est.optimise_pose()
        assert it < 50
        finals.append(est.predicted_pose())
        cov = est.covariance()
        assert np.allclose(cov, cov.T, rtol=1e-9, atol=1e-18) and np.all(np.linalg.eigvalsh((cov + cov.T) / 2) > 0)
    for F in finals[1:]:
        dt, dr = pose_error(F, finals[0])
        assert dt < 1e-9 and dr < 1e-9, (dt, dr)


def test_determine_roi_stays_inside_the_image_and_covers_the_predictions():
    """LEDDetector::determineROI (led_detector.cpp:114-179): the rectangle is clamped to the image, never empty, and contains the
    (re-distorted) predicted LED pixels that lie inside the image with the border around them."""
    rng = np.random.default_rng(12)
    K, D = synth.camera()
    markers = synth.make_cold_scene(1, n_leds=5, seed=1).markers
    params = synth.Params()
    n_inside = 0
    for _ in range(200):
        T = _random_pose(rng, z=(0.3, 2.0))
        T[:3, 3] += rng.normal(size=3) * [0.3, 0.3, 0.0]                 # some objects partly or wholly outside the image
        est = pose_oracle.PoseEstimatorOracle(K, D, markers, params)
        est.set_predicted_pose(T)
        est.L.mpeo_predict_marker_positions(est.h)                       # predictMarkerPositionsInImage (pose_estimator.cpp:270-276)
        x, y, w, h = est.determine_roi(752, 480)
        assert 0 <= x and 0 <= y and w >= 1 and h >= 1 and x + w <= 752 and y + h <= 480, (x, y, w, h)
        _, cam = _project(K, T, markers)
        if cam[:, 2].min() <= 0.1:
            continue
        dist, _, _ = synth.project_distorted(K, D, T, markers)
        inside = (dist[:, 0] > 25) & (dist[:, 0] < 727) & (dist[:, 1] > 25) & (dist[:, 1] < 455)
        if inside.all():
            n_inside += 1
            b = params.roi_border_thickness - 2.0                         # the corners are re-distorted, not the points: allow 2 px
            assert (dist[:, 0] >= x - 2).all() and (dist[:, 0] <= x + w + 2).all() and (dist[:, 1] >= y - 2).all() and (dist[:, 1] <= y + h + 2).all()
            assert w >= np.ptp(dist[:, 0]) + b and h >= np.ptp(dist[:, 1]) + b
    assert n_inside >= 30


def test_find_leds_on_a_roi_equals_find_leds_on_the_cropped_image():
    """LEDDetector::findLeds works on image(ROI) only (led_detector.cpp:44-57: the blur sees a fresh Mat, BORDER_REFLECT_101 at the
    ROI edge): searching an ROI must equal searching a physical copy of that crop and shifting the centres by ROI.tl in float32 —
    the property the GPU tiles rely on when they never read pixels outside the ROI.  And integer shifts of the whole image shift
    the distorted centres by exactly that amount."""
    from tests.helpers import oracle_find_leds, random_blob_image
    rng = np.random.default_rng(3)
    K, D = synth.camera()
    params = synth.Params()
    n_cmp = 0
    for it in range(80):
        img = random_blob_image(rng, 240, 320, n_blobs=14, kind="mixed")
        x, y = int(rng.integers(0, 200)), int(rng.integers(0, 150))
        w, h = int(rng.integers(8, 320 - x + 1)), int(rng.integers(8, 240 - y + 1))
        _, ce_roi = oracle_find_leds(img, (x, y, w, h), params, K, D)
        crop = np.ascontiguousarray(img[y:y + h, x:x + w])
        _, ce_crop = oracle_find_leds(crop, (0, 0, w, h), params, K, D)
        assert len(ce_roi) == len(ce_crop)
        if len(ce_crop):
            shifted = ce_crop + np.array([x, y], np.float32)               # Point2f + Point2f
            assert np.array_equal(ce_roi, shifted.astype(np.float32)), it
            n_cmp += len(ce_crop)
        # shift invariance (content moved by whole pixels inside a larger black frame)
        big = np.zeros((300, 400), np.uint8)
        dx, dy = int(rng.integers(0, 80)), int(rng.integers(0, 60))
        big[dy:dy + 240, dx:dx + 320] = img
        _, ce_a = oracle_find_leds(img, (0, 0, 320, 240), params, K, D)
        _, ce_b = oracle_find_leds(big, (dx, dy, 320, 240), params, K, D)
        assert np.array_equal(ce_b, (ce_a + np.array([dx, dy], np.float32)).astype(np.float32)) if len(ce_a) else len(ce_b) == 0
    assert n_cmp > 20
