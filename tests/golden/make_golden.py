"""Generates tests/golden/*.npz from the CPU oracle in THIS container (cv2 4.13 + oracle/libpose_oracle.so).
The reference ships no golden vectors; these pin the oracle's behaviour so that a change in cv2 /
libstdc++ / glibc on another box, or an accidental edit of the oracle, is detected.  Run: python tests/golden/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rpg_monocular_pose_estimator_b200 import synth
from oracle import pose_oracle, find_leds_cv2

out = os.path.dirname(os.path.abspath(__file__))
for n_leds, n_frames, seed in [(4, 6, 31), (5, 8, 32), (8, 3, 33)]:
    sc = synth.make_cold_scene(n_frames, n_leds=n_leds, seed=seed)
    rec = dict(seed=seed, n_leds=n_leds, frames_crc=np.array([int(np.sum(f.astype(np.uint64) * (np.arange(f.size, dtype=np.uint64).reshape(f.shape) % 65521))) for f in sc.frames], np.uint64))
    dets, cents, hists, corrs, poses, covs, iters, upd, init_pose, ncorr = [], [], [], [], [], [], [], [], [], []
    for f in range(n_frames):
        px, ce = find_leds_cv2.find_leds(sc.frames[f], (0, 0, sc.width, sc.height), sc.params.threshold_value, sc.params.gaussian_sigma,
                                         sc.params.min_blob_area, sc.params.max_blob_area, sc.params.max_width_height_distortion,
                                         sc.params.max_circular_distortion, sc.K, sc.D)
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        est.set_image_points(px)
        ok = est.initialise()
        hists.append(est.histogram()); cpad = np.zeros((n_leds, 2), np.uint32); c_ = est.correspondences(); cpad[:len(c_)] = c_; corrs.append(cpad); ncorr.append(len(c_)); init_pose.append(est.predicted_pose())
        it = est.optimise_pose() if ok else 0
        dets.append(px); cents.append(ce); poses.append(est.predicted_pose()); covs.append(est.covariance()); iters.append(it); upd.append(ok)
    np.savez_compressed(os.path.join(out, f"cold_{n_leds}leds.npz"), det=np.array(dets), centers=np.array(cents), hist=np.array(hists),
                        corr=np.array(corrs), n_corr=np.array(ncorr), pose=np.array(poses), init_pose=np.array(init_pose), cov=np.array(covs), iters=np.array(iters), ok=np.array(upd), **rec)
# P3P known-answer vectors
rng = np.random.default_rng(99)
F, P, S, R = [], [], [], []
for i in range(64):
    pts = rng.uniform(-0.2, 0.2, size=(3, 3))
    Rm = synth.rodrigues(rng.normal(size=3) * 0.8); t = np.array([rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(.4, 1.5)])
    cam = (Rm @ pts.T).T + t; f = cam / np.linalg.norm(cam, axis=1, keepdims=True)
    if i % 5 == 0: f = f + rng.normal(size=f.shape) * 0.05; f /= np.linalg.norm(f, axis=1, keepdims=True)
    rc, sol = pose_oracle.p3p(f.T, pts.T)
    F.append(f.T); P.append(pts.T); S.append(sol); R.append(rc)
np.savez_compressed(os.path.join(out, "p3p_kat.npz"), f=np.array(F), P=np.array(P), sol=np.array(S), rc=np.array(R))
print("golden written")

# Tracking sequences (estimateBodyPose frame by frame: cold start, ROI tracking, a whole-image retry after two blank frames)
def tracking_golden(seed, n_frames=24, blank=(9, 10)):
    sc = synth.make_stream_scene(n_frames, n_leds=5, seed=seed)
    for b in blank:
        sc.frames[b][:] = 0
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    upd, roi, ncorr, corr, pose, iters, ndet = [], [], [], [], [], [], []
    for f in range(n_frames):
        u = est.estimate_body_pose(sc.frames[f], sc.times[f])
        upd.append(u); roi.append(est.region_of_interest); ndet.append(est.n_det)
        c_ = est.correspondences() if u else np.zeros((0, 2), np.uint32)
        cpad = np.zeros((5, 2), np.uint32); cpad[:len(c_)] = c_
        corr.append(cpad); ncorr.append(len(c_)); pose.append(est.predicted_pose()); iters.append(est.gn_iterations() if u else 0)
    return dict(seed=seed, blank=np.array(blank), updated=np.array(upd), roi=np.array(roi), n_det=np.array(ndet), n_corr=np.array(ncorr), corr=np.array(corr),
                pose=np.array(pose), iters=np.array(iters))

np.savez_compressed(os.path.join(out, "tracking_5leds.npz"), **tracking_golden(34))
print("tracking golden written")
