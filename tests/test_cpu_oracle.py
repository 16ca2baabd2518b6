"""CPU-only tests (run everywhere, `-m "not gpu"`): the oracle against its committed golden vectors, the pinned
micro-specs of cv2 (fixed-point Gaussian, contour order / external rule), the index-table order of
combinations.cpp, and the host-side logic that does not need a GPU."""
import ctypes as C
import itertools
import math
import os
import re

import cv2
import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth, _lib
from oracle import pose_oracle
from tests.helpers import oracle_find_leds, pose_error, random_blob_image

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_tracking_scene():
    g = np.load(os.path.join(GOLD, "tracking_5leds.npz"))
    sc = synth.make_stream_scene(len(g["updated"]), n_leds=5, seed=int(g["seed"]))
    for b in g["blank"]:
        sc.frames[int(b)][:] = 0
    return g, sc


def test_oracle_reproduces_golden_tracking_sequence():
    """estimateBodyPose frame by frame (cold start, ROI tracking with predictPose / determineROI / findCorrespondences, two blank
    frames with whole-image retries) against the committed fixture: flags, ROIs, correspondences and iteration counts exactly,
    poses to 1e-9."""
    g, sc = _golden_tracking_scene()
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    for f in range(len(g["updated"])):
        upd = est.estimate_body_pose(sc.frames[f], sc.times[f])
        assert upd == bool(g["updated"][f]), f
        assert tuple(est.region_of_interest) == tuple(g["roi"][f]), f
        assert est.n_det == g["n_det"][f], f
        if upd:
            k = int(g["n_corr"][f])
            assert np.array_equal(est.correspondences(), g["corr"][f][:k]), f
            assert est.gn_iterations() == g["iters"][f], f
            dt, dr = pose_error(est.predicted_pose(), g["pose"][f])
            assert dt < 1e-9 and dr < 1e-9, (f, dt, dr)
    assert g["updated"].sum() == len(g["updated"]) - 2 and g["roi"][1][2] < 752


@pytest.mark.parametrize("n_leds,seed", [(4, 31), (5, 32), (8, 33)])
def test_oracle_reproduces_golden(n_leds, seed):
    g = np.load(os.path.join(GOLD, f"cold_{n_leds}leds.npz"))
    sc = synth.make_cold_scene(len(g["det"]), n_leds=n_leds, seed=seed)
    for f in range(len(g["det"])):
        px, ce = oracle_find_leds(sc.frames[f], (0, 0, sc.width, sc.height), sc.params, sc.K, sc.D)
        assert np.array_equal(ce, g["centers"][f]) and np.array_equal(px, g["det"][f])          # findLeds: bit exact
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        est.set_image_points(px)
        ok = est.initialise()
        assert ok == g["ok"][f]
        k = int(g["n_corr"][f])
        assert np.array_equal(est.histogram(), g["hist"][f]) and np.array_equal(est.correspondences(), g["corr"][f][:k])
        assert np.allclose(est.predicted_pose(), g["init_pose"][f], rtol=0, atol=1e-9)
        if ok:
            assert est.optimise_pose() == g["iters"][f]
            dt, dr = pose_error(est.predicted_pose(), g["pose"][f])
            assert dt < 1e-9 and dr < 1e-9
            dt, dr = pose_error(est.predicted_pose(), sc.poses[f])      # close to the pose the frame was rendered with
            assert dt < 0.02 and dr < 0.05


def test_p3p_known_answers_and_geometry():
    g = np.load(os.path.join(GOLD, "p3p_kat.npz"))
    for i in range(len(g["f"])):
        rc, sol = pose_oracle.p3p(g["f"][i], g["P"][i])
        assert rc == g["rc"][i]
        assert np.array_equal(np.isfinite(sol), np.isfinite(g["sol"][i]))
        m = np.isfinite(sol)
        assert np.allclose(sol[m], g["sol"][i][m], rtol=1e-9, atol=1e-12)
    rng = np.random.default_rng(5)
    for _ in range(50):   # for consistent bearings one of the four solutions reproduces the generating pose
        pts = rng.uniform(-0.2, 0.2, size=(3, 3))
        Rm = synth.rodrigues(rng.normal(size=3) * 0.7); t = np.array([0.1, -0.05, 0.9])
        cam = (Rm @ pts.T).T + t
        f = cam / np.linalg.norm(cam, axis=1, keepdims=True)
        rc, sol = pose_oracle.p3p(f.T, pts.T)
        assert rc == 0
        best = min(np.abs(s[:, :3] - Rm.T).max() + np.abs(s[:, 3] + Rm.T @ t).max() for s in sol if np.isfinite(s).all())
        assert best < 1e-7
    assert pose_oracle.p3p(np.eye(3), np.array([[0, 1, 2.], [0, 0, 0], [0, 0, 0]]))[0] == -1   # colinear world points


def test_quartic_roots():
    r = pose_oracle.solve_quartic([1, -10, 35, -50, 24])          # (x-1)(x-2)(x-3)(x-4)
    assert np.allclose(sorted(r), [1, 2, 3, 4], atol=1e-9)


def _emulate_reference_combinations(N, K=3):
    """Own emulation of the working-vector iteration of Combinations::combinationsNoReplacement
    (monocular_pose_estimator_lib/src/combinations.cpp:60-125): its row order is lexicographic."""
    wv = list(range(1, K + 1)); lim, idx = K, 1
    bc = math.comb(N, K); rows = [list(wv)]
    for i in range(2, bc):
        if idx + lim < N: step, flag = idx, 0
        else: step, flag = 1, 1
        for j in range(1, step + 1): wv[K + j - idx - 1] = lim + j
        rows.append(list(wv)); idx = idx * flag + 1; lim = wv[K - idx]
    rows.append(list(range(N - K + 1, N + 1)))
    return rows


@pytest.mark.parametrize("N", [4, 5, 6, 8, 11])
def test_combination_order_is_lexicographic(N):
    assert _emulate_reference_combinations(N) == [list(c) for c in itertools.combinations(range(1, N + 1), 3)]


def _taps(sigma):
    n = int(np.rint(sigma * 6 + 1)) | 1; n2 = (n - 1) // 2
    vals = [math.exp(-0.5 / (sigma * sigma) * (i - n2) ** 2) for i in range(n2)]
    mul = 1.0 / (2 * sum(vals) + 1.0); out = [0] * n; err = 0.0; tot = 0
    for i in range(n2):
        adj = vals[i] * mul * 256.0 + err; v0 = int(np.rint(adj)); err = adj - v0
        out[i] = out[n - 1 - i] = v0; tot += 2 * v0
    out[n2] = 256 - tot
    return out


def test_gaussian_fixed_point_model_matches_cv2():
    """Pins the blur spec K1a implements: bit-exact kernel taps, 8.8 -> 16.16 fixed point, REFLECT_101."""
    assert _taps(0.6) == [1, 42, 170, 42, 1]
    def refl(i, N):
        if N == 1: return 0
        while i < 0 or i >= N:
            i = -i if i < 0 else 2 * (N - 1) - i
        return i
    rng = np.random.default_rng(1)
    for sigma in [0.3, 0.6, 0.8, 1.0, 1.4]:
        t = np.array(_taps(sigma), np.int64); n = len(t); r = n // 2
        for _ in range(3):
            H, W = int(rng.integers(1, 40)), int(rng.integers(1, 50))
            img = cv2.threshold(rng.integers(0, 256, (H, W), dtype=np.uint8), 140, 255, cv2.THRESH_TOZERO)[1]
            xi = np.array([[refl(x + k - r, W) for k in range(n)] for x in range(W)])
            yi = np.array([[refl(y + k - r, H) for k in range(n)] for y in range(H)])
            h = (img.astype(np.int64)[:, xi] * t).sum(-1)
            v = (h[yi, :] * t[None, :, None]).sum(1)
            assert np.array_equal(((v + 32768) >> 16).astype(np.uint8), cv2.GaussianBlur(img.copy(), (0, 0), sigma, sigmaY=sigma))


def _trace(mask, x0, y0):
    """Python model of K1b's border follower (OpenCV icvFetchContour, outer border) with the raster-order rejection."""
    H, W = mask.shape
    g = lambda x, y: 0 <= x < W and 0 <= y < H and mask[y, x] != 0
    dx = [1, 1, 0, -1, -1, -1, 0, 1]; dy = [0, -1, -1, -1, 0, 1, 1, 1]
    s = 4
    while True:
        s = (s - 1) & 7
        if g(x0 + dx[s], y0 + dy[s]) or s == 4: break
    if s == 4: return [(x0, y0)]
    i1 = (x0 + dx[s], y0 + dy[s]); i3 = (x0, y0); pts = []
    while True:
        while True:
            s = (s + 1) & 7; i4 = (i3[0] + dx[s], i3[1] + dy[s])
            if g(*i4): break
        pts.append(i3)
        if i4 == (x0, y0) and i3 == i1: return pts
        if i4[1] < y0 or (i4[1] == y0 and i4[0] < x0): return None
        i3 = i4; s = (s + 4) & 7


def _inside(poly, px, py):
    c = False
    for i in range(len(poly)):
        (x0, y0), (x1, y1) = poly[i - 1], poly[i]
        if (y0 > py) != (y1 > py) and px < (x0 if y0 == py else x1): c = not c
    return c


def test_contour_model_matches_cv2_external_contours():
    """The rule set K1b implements — candidates = pixels with background W/NW/N/NE; a candidate is a component start iff its
    border never reaches a raster-smaller pixel; a component is reported iff its start is not inside another component's outer
    polygon; output in reverse raster order — reproduces cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) exactly."""
    rng = np.random.default_rng(0)
    n_cont = 0
    for it in range(120):
        H, W = int(rng.integers(6, 40)), int(rng.integers(6, 48))
        if it % 2:
            m = (rng.random((H, W)) < rng.choice([0.2, 0.45, 0.6, 0.75])).astype(np.uint8) * 255
        else:
            m = cv2.GaussianBlur(random_blob_image(rng, H, W, 6, "mixed", noise_max=1), (0, 0), 0.6)
        pad = np.pad(m != 0, 1)
        cand = pad[1:-1, 1:-1] & ~pad[1:-1, :-2] & ~pad[:-2, :-2] & ~pad[:-2, 1:-1] & ~pad[:-2, 2:]
        starts = {}
        for y, x in zip(*np.nonzero(cand)):
            t = _trace(m, int(x), int(y))
            if t is not None: starts[(int(y), int(x))] = t
        ext = [((y, x), t) for (y, x), t in starts.items() if not any(k != (y, x) and _inside(p, x, y) for k, p in starts.items())]
        ext.sort(reverse=True)
        cs, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        assert [t for _, t in ext] == [[tuple(int(v) for v in p[0]) for p in c] for c in cs]
        n_cont += len(cs)
    assert n_cont > 300


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "mpe_b200.h")).read()
    declared = set(re.findall(r"\b(mpe_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert C.sizeof(_lib.MpeResult) == 968


def test_no_gpu_means_loud_failure():
    import rpg_monocular_pose_estimator_b200 as mpe
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mpe.MpeError):
        mpe.Context(0, 1, 752, 480)


def test_host_mirror_small_math_matches_oracle():
    import rpg_monocular_pose_estimator_b200 as mpe
    rng = np.random.default_rng(2)
    K, D = synth.camera()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for _ in range(50):
        T = np.eye(4); T[:3, :3] = synth.rodrigues(rng.normal(size=3) * rng.choice([1e-12, 1e-3, 0.5, 2.0])); T[:3, 3] = rng.normal(size=3) * rng.choice([0, 0.3])
        xi = np.zeros(6); pose_oracle.lib().mpeo_logarithm_map(dp(T), dp(xi))
        assert np.allclose(mpe.PoseEstimator.logarithmMap(T), xi, rtol=1e-9, atol=1e-12)
        E = np.zeros((4, 4)); tw = rng.normal(size=6) * 0.3
        pose_oracle.lib().mpeo_exponential_map(dp(tw), dp(E))
        assert np.allclose(mpe.PoseEstimator.exponentialMap(tw), E, rtol=1e-12, atol=1e-15)
        px = np.ascontiguousarray(rng.uniform(-50, 800, size=(5, 2)))
        est = pose_oracle.PoseEstimatorOracle(K, D, synth.markers(5), synth.Params())
        est.L.mpeo_set_predicted_pixels(est.h, dp(px), 5)
        assert mpe.LEDDetector.determineROI(px, (752, 480), 20, K, D) == est.determine_roi(752, 480)


def test_pose_to_message_matches_scipy_and_node_packing():
    """mpe_pose_to_message (host-only): what MPENode::imageCallback packs (monocular_pose_estimator.cpp:160-190)."""
    from scipy.spatial.transform import Rotation
    L = _lib.load_library()
    rng = np.random.default_rng(4)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for i in range(300):
        rv = rng.normal(size=3)
        rv *= rng.uniform(0, math.pi) / np.linalg.norm(rv)
        if i % 10 == 0:
            rv *= (math.pi - 1e-9) / np.linalg.norm(rv)              # trace near -1: the largest-diagonal branch
        Rm = Rotation.from_rotvec(rv).as_matrix()
        T = np.eye(4); T[:3, :3] = Rm; T[:3, 3] = rng.normal(size=3)
        cov = rng.normal(size=(6, 6))
        pos, q, cm = np.zeros(3), np.zeros(4), np.zeros(36)
        L.mpe_pose_to_message(dp(np.ascontiguousarray(T)), dp(np.ascontiguousarray(cov)), dp(pos), dp(q), dp(cm))
        assert np.array_equal(pos, T[:3, 3])
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.abs(Rotation.from_quat(q).as_matrix() - Rm).max() < 1e-12
        if np.trace(Rm) > 0:
            assert q[3] > 0                                           # Eigen's trace branch: w = sqrt(trace + 1) / 2
        assert np.array_equal(cm.reshape(6, 6), cov)                  # elems[j + 6*i] = cov(i, j)
