"""K1 parity: mpe_find_leds (CUDA) against the cv2 oracle of LEDDetector::findLeds — bit exact:
same number of detections, same order, identical float32 centres and identical undistorted positions."""
import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from tests.helpers import oracle_find_leds, random_blob_image

pytestmark = pytest.mark.gpu


def _compare(ctx, img, roi, params, K, D, tag=""):
    ctx.set_camera(K, D)
    ctx.set_params(params)
    px, centers, flags = ctx.find_leds(img, roi)
    opx, ocenters = oracle_find_leds(img, roi, params, K, D)
    assert flags == 0, f"{tag}: flags {flags}"
    assert len(centers) == len(ocenters), f"{tag}: n_det gpu {len(centers)} oracle {len(ocenters)}"
    if len(ocenters):
        assert np.array_equal(centers.view(np.uint32), ocenters.view(np.uint32)), f"{tag}: centres differ\n{centers}\n{ocenters}"
        assert np.array_equal(px, opx), f"{tag}: undistorted positions differ\n{px - opx}"
    return len(ocenters)


def test_synthetic_led_frames_full_image(gpu_ctx_752):
    for n_leds in (4, 5, 8):
        sc = synth.make_cold_scene(6, n_leds=n_leds, seed=1000 + n_leds)
        for f in range(6):
            n = _compare(gpu_ctx_752, sc.frames[f], (0, 0, sc.width, sc.height), sc.params, sc.K, sc.D, f"leds{n_leds} f{f}")
            assert n == n_leds


@pytest.mark.parametrize("kind", ["ellipse", "mixed"])
def test_random_blobs_full_image(gpu_ctx_752, kind):
    rng = np.random.default_rng(7)
    K, D = synth.camera()
    total = 0
    for it in range(40):
        img = random_blob_image(rng, 480, 752, n_blobs=int(rng.integers(1, 25)), kind=kind)
        p = synth.Params(threshold_value=int(rng.choice([100, 127, 128, 140, 200])), min_blob_area=float(rng.choice([0.5, 10, 30])),
                         max_blob_area=float(rng.choice([200, 1000])), max_width_height_distortion=float(rng.choice([0.5, 0.9])),
                         max_circular_distortion=float(rng.choice([0.5, 0.95])))
        total += _compare(gpu_ctx_752, img, (0, 0, 752, 480), p, K, D, f"{kind} it{it}")
    assert total > 50


def test_roi_subrects_and_border_blobs(gpu_ctx_752):
    rng = np.random.default_rng(11)
    K, D = synth.camera()
    p = synth.Params(min_blob_area=1.0, max_blob_area=2000, max_width_height_distortion=0.95, max_circular_distortion=0.99)
    total = 0
    for it in range(60):
        img = random_blob_image(rng, 480, 752, n_blobs=30, kind="mixed")
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        if it % 10 == 0:
            w, h = int(rng.integers(1, 6)), int(rng.integers(1, 6))       # tiny ROIs: reflect bounces
        x, y = int(rng.integers(0, 752 - w + 1)), int(rng.integers(0, 480 - h + 1))
        total += _compare(gpu_ctx_752, img, (x, y, w, h), p, K, D, f"roi it{it} {(x, y, w, h)}")
    assert total > 30


@pytest.mark.parametrize("sigma", [0.3, 0.45, 0.6, 0.8, 1.0, 1.2, 1.4, 1.5, 2.0, 3.0, 6.0])
def test_other_sigmas(gpu_ctx_752, sigma):
    rng = np.random.default_rng(int(sigma * 100))
    K, D = synth.camera()
    p = synth.Params(gaussian_sigma=sigma, min_blob_area=1.0, max_blob_area=3000, max_width_height_distortion=0.95,
                     max_circular_distortion=0.99)
    for it in range(6):
        img = random_blob_image(rng, 480, 752, n_blobs=12, kind="mixed")
        roi = (0, 0, 752, 480) if it % 2 == 0 else (100, 50, 333, 217)
        _compare(gpu_ctx_752, img, roi, p, K, D, f"sigma{sigma} it{it}")


def test_thresholds_and_bright_background(gpu_ctx_752):
    rng = np.random.default_rng(3)
    K, D = synth.camera()
    for thr in (0, 1, 50, 127, 128, 129, 254, 255):
        img = random_blob_image(rng, 480, 752, n_blobs=10, kind="mixed", noise_max=int(rng.choice([20, 100, 200])))
        p = synth.Params(threshold_value=thr, min_blob_area=1.0, max_blob_area=1e9, max_width_height_distortion=1.0,
                         max_circular_distortion=1.0)
        # a bright background makes one huge component: check it does not break (area filter open)
        if thr < 100:
            img = img[:96, :128].copy()
            import rpg_monocular_pose_estimator_b200 as mpe
        _compare(gpu_ctx_752, img, (0, 0, img.shape[1], img.shape[0]), p, K, D, f"thr{thr}")


def test_empty_and_saturated(gpu_ctx_752):
    K, D = synth.camera()
    p = synth.Params()
    img = np.zeros((480, 752), np.uint8)
    assert _compare(gpu_ctx_752, img, (0, 0, 752, 480), p, K, D, "black") == 0
    img[:] = 255
    _compare(gpu_ctx_752, img, (0, 0, 752, 480), p, K, D, "white")
    _compare(gpu_ctx_752, img, (10, 10, 40, 40), synth.Params(max_blob_area=1e6, max_circular_distortion=1.0), K, D, "white roi")


def test_1080p_two_column_tiles():
    import rpg_monocular_pose_estimator_b200 as mpe
    ctx = mpe.Context(0, 2, 1920, 1080)
    try:
        K, D = synth.camera(1920, 1080)
        sc = synth.make_cold_scene(3, n_leds=5, width=1920, height=1080, seed=77)
        for f in range(3):
            assert _compare(ctx, sc.frames[f], (0, 0, 1920, 1080), sc.params, sc.K, sc.D, f"1080p f{f}") == 5
        rng = np.random.default_rng(5)
        p = synth.Params(min_blob_area=1.0, max_blob_area=5000, max_width_height_distortion=0.95, max_circular_distortion=0.99)
        for it in range(6):
            img = random_blob_image(rng, 1080, 1920, n_blobs=60, kind="mixed")
            # blobs straddling the column-tile seam at x = 960
            for yy in range(50, 1000, 90):
                img[yy:yy + 9, 955:966] = 250
            roi = (0, 0, 1920, 1080) if it % 2 == 0 else (17, 33, 1800, 900)
            _compare(ctx, img, roi, p, K, D, f"1080p random it{it}")
    finally:
        ctx.close()


def test_odd_image_sizes_and_pitch():
    """Widths that are not multiples of 4/16/32 and a padded host pitch: the device copy is pitched to 16 bytes and the
    tensor map sees the padded row; results must not depend on what lies in the padding."""
    import rpg_monocular_pose_estimator_b200 as mpe
    rng = np.random.default_rng(17)
    for (w, h) in [(750, 470), (641, 479), (333, 97), (67, 33)]:
        ctx = mpe.Context(0, 2, w, h)
        try:
            K, D = synth.camera(w, h)
            p = synth.Params(min_blob_area=1.0, max_blob_area=5000, max_width_height_distortion=0.95, max_circular_distortion=0.99)
            for it in range(4):
                big = random_blob_image(rng, h, w + 13, n_blobs=20, kind="mixed")
                img = big[:, :w] if it % 2 else np.ascontiguousarray(big[:, :w])     # odd iterations: non-contiguous rows (pitch = w + 13)
                roi = (0, 0, w, h) if it < 2 else (3, 2, w - 5, h - 3)
                ctx.set_camera(K, D); ctx.set_params(p)
                px, centers, flags = ctx.find_leds_strided(img, roi)
                opx, ocenters = oracle_find_leds(np.ascontiguousarray(img), roi, p, K, D)
                assert flags == 0 and len(centers) == len(ocenters), (w, h, it)
                if len(ocenters):
                    assert np.array_equal(centers.view(np.uint32), ocenters.view(np.uint32)) and np.array_equal(px, opx), (w, h, it)
        finally:
            ctx.close()


def test_error_codes(gpu_ctx_752):
    import rpg_monocular_pose_estimator_b200 as mpe
    K, D = synth.camera()
    gpu_ctx_752.set_camera(K, D)
    with pytest.raises(mpe.MpeError):                      # sigma beyond the largest supported radius
        gpu_ctx_752.set_params(synth.Params(gaussian_sigma=7.0))
    gpu_ctx_752.set_params(synth.Params())
    img = np.zeros((480, 752), np.uint8)
    with pytest.raises(mpe.MpeError):                      # ROI outside the image: cv::Mat::operator() would throw in the reference
        gpu_ctx_752.find_leds(img, (700, 0, 100, 100))
    with pytest.raises(mpe.MpeError):                      # larger than the context was created for
        gpu_ctx_752.find_leds(np.zeros((481, 752), np.uint8), (0, 0, 10, 10))
    with pytest.raises(mpe.MpeError):
        gpu_ctx_752.set_camera(K, np.zeros(3))             # 3 distortion coefficients is not a model OpenCV accepts either
    gpu_ctx_752.set_camera(K, D)


@pytest.mark.parametrize("n_coeffs", [0, 4, 5, 8, 12])
def test_distortion_models_with_4_5_8_12_coefficients(gpu_ctx_752, n_coeffs):
    """cv::undistortPoints (led_detector.cpp:97-98) takes 4, 5, 8 or 12 distortion coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]);
    the undistort epilogue of K1b must match it to the last float bit for each of them — rational model and thin-prism terms included —
    on LED frames (all detections kept) and on random blobs spread over the whole image."""
    rng = np.random.default_rng(300 + n_coeffs)
    K, D5 = synth.camera()
    full = np.concatenate([D5, [0.013, -0.021, 0.0042], [0.0011, -0.0007, 0.0009, 0.0004]])     # k4 k5 k6, s1..s4: small but visible (~0.1-1 px)
    D = full[:n_coeffs].copy()
    sc = synth.make_cold_scene(4, n_leds=5, seed=77)
    for f in range(4):
        assert _compare(gpu_ctx_752, sc.frames[f], (0, 0, 752, 480), sc.params, K, D, f"D{n_coeffs} frame {f}") == 5
    p = synth.Params(min_blob_area=1.0, max_blob_area=2000, max_width_height_distortion=0.95, max_circular_distortion=0.99)
    total = 0
    for it in range(12):
        img = random_blob_image(rng, 480, 752, n_blobs=25, kind="ellipse")
        total += _compare(gpu_ctx_752, img, (0, 0, 752, 480), p, K, D, f"D{n_coeffs} blobs {it}")
    assert total > 60
    if n_coeffs in (8, 12):          # the extra terms really move the points (the test would be vacuous otherwise)
        ctx = gpu_ctx_752
        ctx.set_camera(K, D); ctx.set_params(sc.params)
        a, _, _ = ctx.find_leds(sc.frames[0], (0, 0, 752, 480))
        ctx.set_camera(K, full[:5]); b, _, _ = ctx.find_leds(sc.frames[0], (0, 0, 752, 480))
        assert np.abs(a - b).max() > 1e-3


def test_two_devices_in_one_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: a context on a second device of the same process must get
    its own call (the scan kernel needs ~57 KB of dynamic shared memory at 752x480, above the 48 KB default).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import rpg_monocular_pose_estimator_b200 as mpe
    sc = synth.make_cold_scene(2, n_leds=5, seed=11)
    outs = []
    for dev in (0, 1, 0):
        ctx = mpe.Context(dev, 4, 752, 480)
        outs.append(_compare(ctx, sc.frames[0], (0, 0, 752, 480), sc.params, sc.K, sc.D, f"device {dev}"))
        ctx.set_markers(sc.markers)
        from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
        res = results_to_arrays(ctx.estimate_batch(sc.frames))
        assert res["updated"].sum() == 2
        ctx.close()
    assert outs == [5, 5, 5]
