"""The CUDA path against the committed golden fixtures (tests/golden/*.npz, written by make_golden.py from the CPU oracle) —
no oracle call in these tests: detections bit exact, histogram / correspondences / iteration counts identical, poses within the
north_star tolerance (1e-6 m, 1e-6 rad), P3P known answers to 1e-7 relative with identical NaN patterns."""
import os

import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
from tests.helpers import pose_error

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n_leds,seed", [(4, 31), (5, 32), (8, 33)])
def test_cold_path_matches_golden(gpu_ctx_752, n_leds, seed):
    g = np.load(os.path.join(GOLD, f"cold_{n_leds}leds.npz"))
    n = len(g["det"])
    sc = synth.make_cold_scene(n, n_leds=n_leds, seed=seed)
    ctx = gpu_ctx_752
    ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
    res = results_to_arrays(ctx.estimate_batch(sc.frames))
    for f in range(n):
        # LEDDetector::findLeds: bit exact
        px, ce, _ = ctx.find_leds(sc.frames[f], (0, 0, sc.width, sc.height))
        assert np.array_equal(ce, g["centers"][f]) and np.array_equal(px, g["det"][f]), f
        # PoseEstimator::initialise: histogram and decoded correspondences identical
        ok, hist, corr, pose0 = ctx.initialise(g["det"][f])
        k = int(g["n_corr"][f])
        assert ok == g["ok"][f] and np.array_equal(hist, g["hist"][f]) and np.array_equal(corr, g["corr"][f][:k]), f
        r = res[f]
        assert r["updated"] == g["ok"][f] and r["n_det"] == len(g["det"][f]), f
        if g["ok"][f]:
            dt, dr = pose_error(pose0, g["init_pose"][f])
            assert dt < 1e-6 and dr < 1e-6, (f, dt, dr)
            assert r["n_corr"] == k and np.array_equal(r["corr"][:2 * k].reshape(k, 2), g["corr"][f][:k]), f
            assert r["gn_iters"] == g["iters"][f], f
            dt, dr = pose_error(r["pose"].reshape(4, 4), g["pose"][f])
            assert dt < 1e-6 and dr < 1e-6, (f, dt, dr)
            co = g["cov"][f]
            assert np.allclose(r["cov"].reshape(6, 6), co, rtol=1e-6, atol=1e-12 * np.abs(co).max()), f


def test_p3p_known_answers(gpu_ctx_752):
    g = np.load(os.path.join(GOLD, "p3p_kat.npz"))
    st, sol = gpu_ctx_752.p3p(g["f"], g["P"])
    assert np.array_equal(st, g["rc"])
    for i in range(len(st)):
        if st[i] != 0:
            continue
        a, b = sol[i], g["sol"][i]
        assert np.array_equal(np.isfinite(a), np.isfinite(b)), i
        m = np.isfinite(b)
        if m.any():
            assert (np.abs(a[m] - b[m]) / np.maximum(1.0, np.abs(b[m]))).max() < 1e-7, i
