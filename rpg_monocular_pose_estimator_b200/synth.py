"""Seeded synthetic scenes for tests and bench.

Fixtures taken from the reference repository (inputs only — the reference has no outputs to copy):
  * camera K / D for 752x480:  /root/reference/README.md:165-166
  * demo marker geometry:      /root/reference/monocular_pose_estimator/marker_positions/demo_marker_positions.yaml:4-15
  * demo parameters:           /root/reference/monocular_pose_estimator/launch/demo.launch:12-22

Everything here is numpy on the host; it produces the bytes that are fed identically to the CUDA
path and to the CPU oracle.  It is input generation, not part of the measured hot path.
"""
from __future__ import annotations

import dataclasses
import numpy as np

# README.md:165-166
K_752 = np.array([[615.652408400557, 0.0, 362.655454167686],
                  [0.0, 616.760184718123, 256.67210750994],
                  [0.0, 0.0, 1.0]], dtype=np.float64)
D_752 = np.array([-0.358561237166698, 0.149312912580924, 0.000484551782515636,
                  -0.000200189442379448, 0.0], dtype=np.float64)

# demo_marker_positions.yaml:4-15
MARKERS_4 = np.array([[0.0714197, 0.0800214, 0.0622611],
                      [0.0400755, -0.0912328, 0.0317064],
                      [-0.0647293, -0.0879977, 0.0830852],
                      [-0.0558663, -0.0165446, 0.053473]], dtype=np.float64)
# 5th LED: one extra non-coplanar, non-symmetric point inside (-0.1,0.1)^3 m, chosen once and frozen.
MARKERS_5 = np.vstack([MARKERS_4, [[0.0213846, 0.0362517, -0.0748301]]])
# 8 LEDs: eight points inside (-0.2,0.2)^3 m, chosen once and frozen (half-extent 0.2 m).
MARKERS_8 = np.array([[-0.1090735, 0.1372027, 0.1127286],
                      [0.0741443, -0.1564329, 0.1889113],
                      [0.1613262, -0.0395441, -0.1420712],
                      [-0.1731148, -0.1206042, -0.0538374],
                      [0.0318327, 0.1816411, -0.1695229],
                      [0.1927415, 0.1208643, 0.0654419],
                      [-0.0455871, -0.0271365, 0.1936642],
                      [-0.1386523, 0.0642918, -0.1817350]], dtype=np.float64)


def markers(n_leds: int) -> np.ndarray:
    return {4: MARKERS_4, 5: MARKERS_5, 8: MARKERS_8}[n_leds].copy()


@dataclasses.dataclass
class Params:
    """The 11 dynamic-reconfigure tunables (cfg/MonocularPoseEstimator.cfg:12-22), demo.launch values."""
    threshold_value: int = 140
    gaussian_sigma: float = 0.6
    min_blob_area: float = 10.0
    max_blob_area: float = 200.0
    max_width_height_distortion: float = 0.5
    max_circular_distortion: float = 0.5
    back_projection_pixel_tolerance: float = 5.0
    nearest_neighbour_pixel_tolerance: float = 7.0
    certainty_threshold: float = 0.75
    valid_correspondence_threshold: float = 0.7
    roi_border_thickness: int = 20


def camera(width: int = 752, height: int = 480):
    """K, D for the given resolution (1080p: intrinsics scaled, same D)."""
    K = K_752.copy()
    sx, sy = width / 752.0, height / 480.0
    K[0, 0] *= sx; K[0, 2] *= sx
    K[1, 1] *= sy; K[1, 2] *= sy
    return K, D_752.copy()


def rodrigues(w: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(w))
    if th == 0.0:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def project_distorted(K, D, T, pts_obj):
    """Object points -> distorted pixel coordinates (plumb-bob forward model) and undistorted pixels."""
    P = (T[:3, :3] @ pts_obj.T).T + T[:3, 3]
    x = P[:, 0] / P[:, 2]
    y = P[:, 1] / P[:, 2]
    und = np.stack([K[0, 0] * x + K[0, 1] * y + K[0, 2], K[1, 1] * y + K[1, 2]], axis=1)
    k1, k2, p1, p2, k3 = D[:5]
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    dist = np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], axis=1)
    return dist, und, P[:, 2]


def sample_pose(rng, K, D, pts_obj, width, height, z_range=(0.4, 1.2), margin=30.0, min_sep=14.0,
                max_rot=0.7, max_tries=1000):
    """A random object pose whose LEDs all land inside the image with margin and do not merge."""
    for _ in range(max_tries):
        z = rng.uniform(*z_range)
        # choose a pixel for the object origin, back-project at depth z
        u = rng.uniform(margin + 40, width - margin - 40)
        v = rng.uniform(margin + 40, height - margin - 40)
        t = np.array([(u - K[0, 2]) / K[0, 0] * z, (v - K[1, 2]) / K[1, 1] * z, z])
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * rng.uniform(0, max_rot)
        T = np.eye(4)
        T[:3, :3] = rodrigues(w)
        T[:3, 3] = t
        dist, _, depth = project_distorted(K, D, T, pts_obj)
        if np.any(depth < 0.2):
            continue
        if (dist[:, 0].min() < margin or dist[:, 0].max() > width - margin or
                dist[:, 1].min() < margin or dist[:, 1].max() > height - margin):
            continue
        d = np.linalg.norm(dist[:, None, :] - dist[None, :, :], axis=-1) + np.eye(len(dist)) * 1e9
        if d.min() < min_sep:
            continue
        return T
    raise RuntimeError("could not sample a valid pose")


def render_frame(rng, width, height, led_px, spot_sigma=2.0, amplitude=400.0, out=None):
    """u8 frame: background noise 0..19 plus one saturating Gaussian spot per LED."""
    if out is None:
        out = np.empty((height, width), np.uint8)
    out[...] = rng.integers(0, 20, size=(height, width), dtype=np.uint8)
    r = int(np.ceil(4 * spot_sigma)) + 1
    for (cx, cy) in led_px:
        x0, y0 = int(np.floor(cx)) - r, int(np.floor(cy)) - r
        xs = np.arange(max(x0, 0), min(x0 + 2 * r + 2, width))
        ys = np.arange(max(y0, 0), min(y0 + 2 * r + 2, height))
        if len(xs) == 0 or len(ys) == 0:
            continue
        g = amplitude * np.exp(-((xs[None, :] - cx) ** 2 + (ys[:, None] - cy) ** 2) / (2 * spot_sigma ** 2))
        patch = out[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1].astype(np.float64) + g
        out[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1] = np.clip(np.rint(patch), 0, 255).astype(np.uint8)
    return out


@dataclasses.dataclass
class Scene:
    width: int
    height: int
    K: np.ndarray
    D: np.ndarray
    markers: np.ndarray          # n x 3
    params: Params
    frames: np.ndarray           # B x H x W u8
    poses: np.ndarray            # B x 4 x 4 ground-truth camera<-object transforms
    times: np.ndarray            # B timestamps (1/60 s apart)


def make_cold_scene(n_frames: int, n_leds: int = 5, width: int = 752, height: int = 480, seed: int = 0,
                    z_range=None, params: Params | None = None) -> Scene:
    """Independent frames (one random pose each): the 'cold' workload — every frame runs the full
    findLeds + initialise + checkCorrespondences + optimisePose path.  frame seed = seed + index."""
    K, D = camera(width, height)
    pts = markers(n_leds)
    if z_range is None:
        z_range = (0.8, 1.0) if n_leds == 8 else (0.4, 1.2)
    frames = np.empty((n_frames, height, width), np.uint8)
    poses = np.empty((n_frames, 4, 4))
    for f in range(n_frames):
        rng = np.random.default_rng(seed + f)
        T = sample_pose(rng, K, D, pts, width, height, z_range=z_range)
        dist, _, _ = project_distorted(K, D, T, pts)
        render_frame(rng, width, height, dist, out=frames[f])
        poses[f] = T
    return Scene(width, height, K, D, pts, params or Params(), frames, poses, np.arange(n_frames) / 60.0)


def make_stream_scene(n_frames: int, n_leds: int = 5, width: int = 752, height: int = 480, seed: int = 0,
                      params: Params | None = None) -> Scene:
    """One smooth trajectory (constant body twist plus a slow drift) so that tracking mode engages."""
    K, D = camera(width, height)
    pts = markers(n_leds)
    rng0 = np.random.default_rng(seed)
    T = sample_pose(rng0, K, D, pts, width, height, z_range=(0.6, 0.9), margin=120.0)
    lin = rng0.normal(size=3) * 0.02 / 60.0      # ~2 cm/s
    ang = rng0.normal(size=3) * 0.15 / 60.0      # ~0.15 rad/s
    frames = np.empty((n_frames, height, width), np.uint8)
    poses = np.empty((n_frames, 4, 4))
    for f in range(n_frames):
        rng = np.random.default_rng(seed + 1000003 * (f + 1))
        dist, _, _ = project_distorted(K, D, T, pts)
        render_frame(rng, width, height, dist, out=frames[f])
        poses[f] = T
        dT = np.eye(4)
        dT[:3, :3] = rodrigues(ang)
        dT[:3, 3] = lin
        T = T @ dT
        # turn around before leaving the image
        dist_next, _, _ = project_distorted(K, D, T, pts)
        if (dist_next[:, 0].min() < 60 or dist_next[:, 0].max() > width - 60 or
                dist_next[:, 1].min() < 60 or dist_next[:, 1].max() > height - 60):
            lin, ang = -lin, -ang
    return Scene(width, height, K, D, pts, params or Params(), frames, poses, np.arange(n_frames) / 60.0)
