"""Device-resident tracking loop (mpe_streams_step_device): S independent streams advanced frame by frame on the GPU must
reproduce, stream by stream and frame by frame, the oracle's estimateBodyPose (pose_estimator.cpp:62-147): update flag,
region of interest, correspondences, Gauss-Newton iteration count, pose within 1e-6 m / 1e-6 rad."""
import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
from oracle import pose_oracle
from tests.helpers import pose_error

pytestmark = pytest.mark.gpu


def _run(streams, ctx, torch):
    S, T = len(streams), len(streams[0].frames)
    sc0 = streams[0]
    H, W = sc0.height, sc0.width
    frames = np.stack([np.stack([streams[s].frames[t] for s in range(S)]) for t in range(T)])      # T x S x H x W
    dev = torch.from_numpy(frames).cuda()
    ctx.set_camera(sc0.K, sc0.D); ctx.set_params(sc0.params); ctx.set_markers(sc0.markers)
    ctx.streams_reset(S)
    ctx.streams_set_frame_map(0, 0)
    out = []
    for t in range(T):
        res = ctx.streams_step_device(dev[t].data_ptr(), W, W * H, W, H, [streams[s].times[t] for s in range(S)])
        out.append(results_to_arrays(res).copy())
    return out


def test_streams_match_oracle_frame_by_frame(gpu_ctx_752):
    import torch
    T = 22
    streams = [synth.make_stream_scene(T, n_leds=5, seed=300 + s) for s in range(6)]
    # stream 1: the LEDs vanish for two frames (ROI search fails, whole-image retry fails), then come back
    streams[1].frames[9][:] = 0
    streams[1].frames[10][:] = 0
    # stream 2: one LED is occluded for a while (4 detections, NN correspondences still fine)
    for t in range(6, 12):
        px, _, _ = synth.project_distorted(streams[2].K, streams[2].D, streams[2].poses[t], streams[2].markers)
        x, y = int(px[0][0]), int(px[0][1])
        streams[2].frames[t][max(y - 12, 0):y + 12, max(x - 12, 0):x + 12] = 0
    # stream 3: a jump (frames of another trajectory spliced in) breaks the prediction: NN check fails -> re-initialise
    other = synth.make_stream_scene(T, n_leds=5, seed=999)
    for t in range(12, T):
        streams[3].frames[t] = other.frames[t]
    out = _run(streams, gpu_ctx_752, torch)
    n_upd = n_retry = n_reinit = 0
    for s, sc in enumerate(streams):
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        for t in range(T):
            upd = est.estimate_body_pose(sc.frames[t], sc.times[t])
            r = out[t][s]
            tag = f"stream {s} frame {t}"
            assert bool(r["updated"]) == upd, tag
            assert tuple(r["roi"]) == tuple(est.region_of_interest), (tag, tuple(r["roi"]), est.region_of_interest)
            n_retry += int(bool(r["flags"] & 16)); n_reinit += int(bool(r["flags"] & 8) and t > 0)
            if upd:
                n_upd += 1
                k = r["n_corr"]
                assert np.array_equal(r["corr"][:2 * k].reshape(k, 2), est.correspondences()), tag
                dt, dr = pose_error(r["pose"].reshape(4, 4), est.predicted_pose())
                assert dt < 1e-6 and dr < 1e-6, (tag, dt, dr)
                assert r["gn_iters"] == est.gn_iterations(), tag
                co = est.covariance()
                assert np.allclose(r["cov"].reshape(6, 6), co, rtol=1e-6, atol=1e-12 * np.abs(co).max()), tag
    assert n_upd >= 0.8 * len(streams) * T
    assert n_retry >= 2          # the blanked frames forced whole-image retries
    assert n_reinit >= 1         # the spliced trajectory forced a brute-force re-initialisation while tracking


def test_streams_with_frame_map_replay(gpu_ctx_752):
    """Many streams replaying a few recorded sequences through the frame map give identical results per replica."""
    import torch
    T, Sd, rep = 8, 3, 4
    seqs = [synth.make_stream_scene(T, n_leds=5, seed=700 + s) for s in range(Sd)]
    H, W = seqs[0].height, seqs[0].width
    buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).cuda()               # (Sd*T) x H x W
    ctx = gpu_ctx_752
    ctx.set_camera(seqs[0].K, seqs[0].D); ctx.set_params(seqs[0].params); ctx.set_markers(seqs[0].markers)
    S = Sd * rep
    ctx.streams_reset(S)
    for t in range(T):
        fmap = torch.tensor([(s % Sd) * T + t for s in range(S)], dtype=torch.int32, device="cuda")
        ctx.streams_set_frame_map(fmap.data_ptr(), Sd * T)
        res = results_to_arrays(ctx.streams_step_device(buf.data_ptr(), W, W * H, W, H, [seqs[s % Sd].times[t] for s in range(S)]))
        for s in range(Sd, S):
            a, b = res[s], res[s % Sd]
            assert a["updated"] == b["updated"] and np.array_equal(a["pose"], b["pose"]) and tuple(a["roi"]) == tuple(b["roi"])
    ctx.streams_set_frame_map(0, 0)
    assert res["updated"].sum() == S


def _edited_streams(T):
    streams = [synth.make_stream_scene(T, n_leds=5, seed=300 + s) for s in range(4)]
    streams[1].frames[9][:] = 0                      # LEDs vanish: ROI search and the whole-image retry fail
    streams[1].frames[10][:] = 0
    other = synth.make_stream_scene(T, n_leds=5, seed=999)
    for t in range(12, T):                           # spliced trajectory: NN check fails -> brute-force re-initialisation
        streams[3].frames[t] = other.frames[t]
    return streams


def test_host_step_single_camera_graph_replay_matches_oracle(gpu_ctx_752):
    """mpe_streams_step with ONE host image per call — the drop-in estimateBodyPose of a camera driver — replayed as a CUDA
    graph from the third frame on, equals the oracle frame by frame (including the retry / re-initialisation ladders)."""
    T = 22
    ctx = gpu_ctx_752
    for sc in (_edited_streams(T)[1], _edited_streams(T)[3], _edited_streams(T)[0]):
        ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
        ctx.streams_reset(1)
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        l0 = ctx.launch_count()
        n_upd = 0
        for t in range(T):
            r = results_to_arrays(ctx.streams_step(sc.frames[t][None], [sc.times[t]]))[0]
            upd = est.estimate_body_pose(sc.frames[t], sc.times[t])
            tag = f"frame {t}"
            assert bool(r["updated"]) == upd, tag
            assert tuple(r["roi"]) == tuple(est.region_of_interest), tag
            if upd:
                n_upd += 1
                k = r["n_corr"]
                assert np.array_equal(r["corr"][:2 * k].reshape(k, 2), est.correspondences()), tag
                dt, dr = pose_error(r["pose"].reshape(4, 4), est.predicted_pose())
                assert dt < 1e-6 and dr < 1e-6, (tag, dt, dr)
                assert r["gn_iters"] == est.gn_iterations(), tag
        assert n_upd >= T - 4
        assert ctx.launch_count() - l0 >= 9 * T       # every step, replayed or not, accounts for its kernel launches (short step: 10, complete step: 21)


def test_host_step_equals_device_step_and_follows_reconfiguration(gpu_ctx_752):
    """Host-image steps (graph replay) and device-buffer steps (plain launches) give bit-identical records; a configuration
    change between two steps invalidates the captured graph."""
    import torch
    T = 14
    streams = _edited_streams(T)
    S = len(streams)
    sc0 = streams[0]
    ctx = gpu_ctx_752
    dev_out = _run(streams, ctx, torch)
    ctx.streams_reset(S)
    for t in range(T):
        frames = np.stack([streams[s].frames[t] for s in range(S)])
        res = results_to_arrays(ctx.streams_step(frames, [streams[s].times[t] for s in range(S)]))
        assert res.tobytes() == dev_out[t].tobytes(), f"frame {t}"
    # reconfigure: threshold 255 -> nothing detected -> no update; the stale graph must not be replayed
    import dataclasses
    p = sc0.params
    p255 = dataclasses.replace(p, threshold_value=255)
    ctx.set_params(p255)
    frames = np.stack([streams[s].frames[T - 1] for s in range(S)])
    res = results_to_arrays(ctx.streams_step(frames, [streams[s].times[T - 1] + 1 / 60 for s in range(S)]))
    assert res["updated"].sum() == 0 and res["n_det"].sum() == 0
    ctx.set_params(p)
    ctx.set_graph_replay(False)
    res = results_to_arrays(ctx.streams_step(frames, [streams[s].times[T - 1] + 2 / 60 for s in range(S)]))
    ctx.set_graph_replay(True)
    assert res["n_det"].min() >= 4


def test_zero_copy_ingest_reads_pinned_images_in_place(gpu_ctx_752):
    """mpe_streams_step with page-locked images: AUTO switches from bulk copies to in-place ROI reads (TMA from host memory)
    once every stream is tracking; the records are bit-identical to the copy mode; ZERO_COPY refuses pageable memory."""
    import torch
    from rpg_monocular_pose_estimator_b200 import MpeError
    T = 10
    streams = [synth.make_stream_scene(T, n_leds=5, seed=300 + s) for s in range(4)]
    S = len(streams)
    sc0 = streams[0]
    ctx = gpu_ctx_752
    ctx.set_camera(sc0.K, sc0.D); ctx.set_params(sc0.params); ctx.set_markers(sc0.markers)
    pinned = torch.from_numpy(np.stack([np.stack([streams[s].frames[t] for s in range(S)]) for t in range(T)])).pin_memory().numpy()
    times = [[streams[s].times[t] for s in range(S)] for t in range(T)]
    out = {}
    for mode in (ctx.INGEST_COPY, ctx.INGEST_AUTO, ctx.INGEST_ZERO_COPY):
        ctx.set_ingest_mode(mode)
        ctx.streams_reset(S)
        st0 = ctx.ingest_stats()
        out[mode] = [results_to_arrays(ctx.streams_step(pinned[t], times[t])).copy() for t in range(T)]
        st1 = ctx.ingest_stats()
        dz, dc = st1["zero_copy_steps"] - st0["zero_copy_steps"], st1["copy_steps"] - st0["copy_steps"]
        if mode == ctx.INGEST_COPY:
            assert (dz, dc) == (0, T)
        elif mode == ctx.INGEST_AUTO:
            assert dc >= 1 and dz >= T - 3, (dz, dc)           # cold start by copy, tracking in place
        else:
            assert (dz, dc) == (T, 0)
    for t in range(T):
        assert out[ctx.INGEST_COPY][t].tobytes() == out[ctx.INGEST_AUTO][t].tobytes() == out[ctx.INGEST_ZERO_COPY][t].tobytes(), t
    assert out[ctx.INGEST_COPY][-1]["updated"].sum() == S
    # pageable images: AUTO falls back to copying, ZERO_COPY is an error
    pageable = np.stack([streams[s].frames[0] for s in range(S)])
    ctx.set_ingest_mode(ctx.INGEST_ZERO_COPY)
    ctx.streams_reset(S)
    with pytest.raises(MpeError):
        ctx.streams_step(pageable, times[0])
    ctx.set_ingest_mode(ctx.INGEST_AUTO)
    r = results_to_arrays(ctx.streams_step(pageable, times[0]))
    assert r["updated"].sum() == S


def test_streams_1080p_match_oracle():
    """Tracking at 1920x1080 (34 strips x 8 column tiles per frame in the K1 tile work list, ROIs that straddle column tiles)."""
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    T, S = 8, 3
    streams = [synth.make_stream_scene(T, n_leds=5, width=1920, height=1080, seed=4100 + s) for s in range(S)]
    ctx = mpe.Context(0, S, 1920, 1080)
    try:
        out = _run(streams, ctx, torch)
        n_upd = 0
        for s, sc in enumerate(streams):
            est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
            for t in range(T):
                upd = est.estimate_body_pose(sc.frames[t], sc.times[t])
                r = out[t][s]
                assert bool(r["updated"]) == upd, (s, t)
                assert tuple(r["roi"]) == tuple(est.region_of_interest), (s, t)
                if upd:
                    n_upd += 1
                    k = r["n_corr"]
                    assert np.array_equal(r["corr"][:2 * k].reshape(k, 2), est.correspondences()), (s, t)
                    dt, dr = pose_error(r["pose"].reshape(4, 4), est.predicted_pose())
                    assert dt < 1e-6 and dr < 1e-6, (s, t, dt, dr)
                    assert r["gn_iters"] == est.gn_iterations(), (s, t)
        assert n_upd >= S * T - 3
        assert out[-1]["roi"][:, 2].max() < 1920       # tracking inside ROIs
    finally:
        ctx.close()


def test_device_loop_matches_golden_tracking_fixture(gpu_ctx_752):
    """The device-resident loop against the committed golden fixture directly (no oracle in the loop): tests/golden/tracking_5leds.npz."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracking_5leds.npz"))
    sc = synth.make_stream_scene(len(g["updated"]), n_leds=5, seed=int(g["seed"]))
    for b in g["blank"]:
        sc.frames[int(b)][:] = 0
    ctx = gpu_ctx_752
    ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
    ctx.streams_reset(1)
    for f in range(len(g["updated"])):
        r = results_to_arrays(ctx.streams_step(sc.frames[f][None], [sc.times[f]]))[0]
        assert bool(r["updated"]) == bool(g["updated"][f]), f
        assert tuple(r["roi"]) == tuple(g["roi"][f]), f
        if r["updated"]:
            assert r["n_det"] == g["n_det"][f], f     # (on frames without detections the fixture holds the size of the stale image_points_)
            k = int(g["n_corr"][f])
            assert r["n_corr"] == k and np.array_equal(r["corr"][:2 * k].reshape(k, 2), g["corr"][f][:k]), f
            assert r["gn_iters"] == g["iters"][f], f
            dt, dr = pose_error(r["pose"].reshape(4, 4), g["pose"][f])
            assert dt < 1e-6 and dr < 1e-6, (f, dt, dr)


def test_more_than_16_detections_in_the_roi_keep_tracking(gpu_ctx_752):
    """Capacity divergence made explicit.  The reference accepts any number of detections; the tables of the brute-force sweep hold
    MPE_MAX_DET = 16.  In TRACKING mode that does not stop the pose path: the nearest neighbours are searched among all (up to 64)
    detections as the reference does, the matched ones are compacted (MPE_F_TOO_MANY_DET is set, correspondence rows refer to the
    compacted list) and checkCorrespondences / optimisePose continue — the pose must be the oracle's, which sees all detections.
    In COLD mode (initialise() with more than 16 detections) the frame is flagged and left without a pose, no error."""
    import torch
    T = 14
    sc = synth.make_stream_scene(T, n_leds=5, seed=4242)
    frames = sc.frames.copy()
    rng = np.random.default_rng(8)
    cluttered = range(6, 11)
    for t in cluttered:                                   # 14 extra spots inside the predicted ROI, >= 12 px away from every LED
        led, _, _ = synth.project_distorted(sc.K, sc.D, sc.poses[t], sc.markers)
        x0, y0 = led.min(0) - 15; x1, y1 = led.max(0) + 15
        spots = []
        while len(spots) < 14:
            p = np.array([rng.uniform(x0, x1), rng.uniform(y0, y1)])
            if min(np.linalg.norm(led - p, axis=1).min(), min([np.linalg.norm(q - p) for q in spots], default=99)) >= 12:
                spots.append(p)
        img = frames[t].astype(np.float64)
        yy, xx = np.mgrid[0:sc.height, 0:sc.width]
        for p in spots:
            img += 400.0 * np.exp(-((xx - p[0]) ** 2 + (yy - p[1]) ** 2) / (2 * 2.0 ** 2))
        frames[t] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    ctx = gpu_ctx_752
    ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
    ctx.streams_reset(1)
    ctx.streams_set_frame_map(0, 0)
    dev = torch.from_numpy(frames).cuda()
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    n_flagged = 0
    for t in range(T):
        r = results_to_arrays(ctx.streams_step_device(dev[t].data_ptr(), sc.width, sc.width * sc.height, sc.width, sc.height, [sc.times[t]]))[0]
        upd = est.estimate_body_pose(frames[t], sc.times[t])
        assert bool(r["updated"]) == upd and upd, t
        dt, dr = pose_error(r["pose"].reshape(4, 4), est.predicted_pose())
        assert dt < 1e-6 and dr < 1e-6 and r["gn_iters"] == est.gn_iterations(), (t, dt, dr)
        if t in cluttered:
            assert est.n_det > 16 and (r["flags"] & 4), (t, est.n_det, r["flags"])
            n_flagged += 1
            k = r["n_corr"]
            corr, ocorr = r["corr"][:2 * k].reshape(k, 2), est.correspondences()
            assert np.array_equal(corr[:, 0], ocorr[:, 0])                                   # same LEDs matched
            odet = est.L.mpeo_get_image_vectors                                                 # (oracle keeps all detections)
            assert r["n_det"] == len(set(ocorr[:, 1].tolist()))                              # compacted to the matched detections
        else:
            assert not (r["flags"] & 4)
    assert n_flagged == len(cluttered)
    # cold mode: a cluttered frame through the batch entry is flagged and has no pose
    res = results_to_arrays(ctx.estimate_batch(frames[8:9]))[0]
    assert res["updated"] == 0 and (res["flags"] & 4) and res["n_det"] > 16
