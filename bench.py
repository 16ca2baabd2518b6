#!/usr/bin/env python
"""bench.py — frames/sec of the per-frame hot path (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[1]): 752x480 synthetic stream, 5 LEDs, full pipeline in COLD mode — every frame
runs whole-image findLeds + initialise (600 P3P solves, 2400 hypotheses) + checkCorrespondences + optimisePose.
One "step" = one pass of that path over one batch of `--batch` frames per GPU.

  value   frames/s with the batch already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e     the same metric through the C-ABI call a user makes (mpe_estimate_batch) with HOST (pinned) frames:
          H2D of every frame and D2H of every result record inside the timed region
  roofline  the HBM-bound kernel (find_leds): algorithmic bytes = W*H per frame / its CUDA-event duration,
            against MEASURED_PEAKS.json; `kernels` lists every kernel's share of the step
  cpu_baseline  the CPU oracle (cv2 findLeds + C++ pose restatement), one thread, bounded sample, same frames

`--impl reference` times the reference's CPU path (the oracle port; the original C++ cannot be built here) on all
host cores instead.  Multi-GPU: one process per GPU (torchrun), frames sharded, no data-path collective; the pose
records are all-gathered over NCCL once per step.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, N_LEDS = 752, 480, 5


def load_traffic(batch, width, height):
    """dram__bytes_read + dram__bytes_write of the scan kernel from the committed `ncu --set full` capture of the same launch
    shape (profiles/ncu_traffic.json, written by profiles/summarise.py); None when no capture of this shape is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if (d.get("batch"), d.get("width"), d.get("height")) == (batch, width, height):
            return d["dram_bytes"]
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t_begin=None, t_end=None):
        """Summarises the samples taken between two wall-clock times (time.time()); nvidia-smi is started long before the
        timed region because its first sample takes about a second."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and (ts < t_begin - 0.02 or ts > t_end + 0.02):
                    continue
                sm_v, mx_v = float(parts[2]), float(parts[3])
            except ValueError:
                continue
            sm.append(sm_v); mx.append(mx_v)
            for n, v in zip(names, parts[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            out["error"] = "no nvidia-smi samples"
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------ CPU arms
def _cpu_worker_init(K, D, markers, params):
    import cv2
    cv2.setNumThreads(1)
    global _W
    from oracle import pose_oracle
    _W = dict(K=K, D=D, markers=markers, params=params, po=pose_oracle)


_FRAMES = None   # set in the parent before the fork so that workers get the frames without any IPC


def _cpu_worker_run(idx):
    po = _W["po"]
    n_upd = 0
    for i in idx:
        fr = _FRAMES[i % len(_FRAMES)]
        est = po.PoseEstimatorOracle(_W["K"], _W["D"], _W["markers"], _W["params"])
        n_upd += int(est.estimate_body_pose(fr, 0.0))
    return n_upd


def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota (the GPU boxes expose 128
    logical CPUs but grant a 16-CPU quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_single_thread(scene, max_seconds=12.0, max_frames=20000):
    import cv2
    cv2.setNumThreads(1)
    from oracle import pose_oracle
    n = 0
    # warm-up
    for f in range(min(5, len(scene.frames))):
        pose_oracle.PoseEstimatorOracle(scene.K, scene.D, scene.markers, scene.params).estimate_body_pose(scene.frames[f], 0.0)
    t0 = time.perf_counter()
    while n < max_frames and time.perf_counter() - t0 < max_seconds:
        est = pose_oracle.PoseEstimatorOracle(scene.K, scene.D, scene.markers, scene.params)
        est.estimate_body_pose(scene.frames[n % len(scene.frames)], 0.0)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference_arm(args):
    """Reference CPU implementation of the path (oracle port: cv2 4.13 findLeds + C++ restatement of the pose code,
    the original cannot be compiled in this image) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from rpg_monocular_pose_estimator_b200 import synth
    cores = usable_cores()
    per_core = 48
    sample = cores * per_core
    scene = synth.make_cold_scene(min(sample, 512), n_leds=N_LEDS, width=WIDTH, height=HEIGHT, seed=args.seed)
    global _FRAMES
    _FRAMES = scene.frames
    chunks = [list(range(i, sample, cores)) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(scene.K, scene.D, scene.markers, scene.params)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker_run, chunks)
        t0 = time.perf_counter()
        upd = 0
        for _ in range(args.steps):
            upd += sum(pool.map(_cpu_worker_run, chunks))
        dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": "frames/sec (752x480, 5 LEDs, cold full pipeline)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
        "config": {"workload": "752x480 synthetic stream, 5 LEDs, cold mode (findLeds + initialise + check + optimisePose per frame)",
                   "frames_per_step": sample, "frames_updated_per_step": upd // args.steps},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frames per step x {args.steps} steps, {cores} processes (cv2 4.13 single-threaded each + C++ oracle)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ tracking mode (secondary)
def run_tracking(args):
    """Steady-state tracking throughput: S streams, each an independent PoseEstimator whose
    state lives on the GPU; one step = one frame for every stream (mpe_streams_step_device).  Sd distinct synthetic
    trajectories of T frames are replayed by S = Sd * rep streams through the frame map; per step Sd distinct frames
    (> L2 for the default sizes) are read.  Single GPU; prints one JSON line (not the headline metric)."""
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth
    from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    W, H = args.width, args.height
    Sd, T = 512, 8 + args.warmup + args.steps
    S = args.batch
    seqs = [synth.make_stream_scene(T, n_leds=args.leds, width=W, height=H, seed=args.seed + 17 * s) for s in range(Sd)]
    buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).to(dev)          # (Sd*T) x H x W
    ctx = mpe.Context(0, S, W, H)
    ctx.set_camera(seqs[0].K, seqs[0].D); ctx.set_params(seqs[0].params); ctx.set_markers(seqs[0].markers)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    ctx.streams_reset(S)
    base = (torch.arange(S, dtype=torch.int32, device=dev) % Sd) * T
    fmap = base.clone()                       # one buffer, advanced in place: a stable key lets the step replay as a CUDA graph
    times = [np.full(S, t / 60.0) for t in range(T)]
    ctx.streams_set_frame_map(fmap.data_ptr(), Sd * T)
    def step(t, fetch=False):
        fmap.copy_(base + t)
        return ctx.streams_step_device(buf.data_ptr(), W, W * H, W, H, times[t], fetch=fetch)
    t = 0
    for _ in range(8 + args.warmup):          # cold start + settle into tracking (it_since_initialized_ == 2), untimed
        step(t); t += 1
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(t); t += 1
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res = results_to_arrays(ctx.fetch_results(S))
    launches = ctx.launch_count() - l0

    # ---- e2e: the same streams fed from a PINNED HOST ring (zero-copy ingest): the findLeds kernels read their ROI tiles in place
    # over PCIe, every step ends with the D2H copy of all result records and a synchronise (wall clock and CUDA events, the larger)
    e2e = None
    if not args.no_e2e:
        host_buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).pin_memory()
        ctx.streams_reset(S)
        t = 0
        def hstep(t):
            fmap.copy_(base + t)
            return ctx.streams_step_device(host_buf.data_ptr(), W, W * H, W, H, times[t], fetch=True)
        for _ in range(8 + args.warmup):
            hstep(t); t += 1
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        w0 = time.perf_counter()
        for _ in range(args.steps):
            r = hstep(t); t += 1
        h1.record(stream)
        torch.cuda.synchronize()
        ms_e2e = max((time.perf_counter() - w0) * 1e3, h0.elapsed_time(h1)) / args.steps
        rr = results_to_arrays(r)
        # bytes that cross PCIe per step: whole TMA boxes of the tiles each ROI touches (32-row strips + 2R halo rows, 256-px column
        # tiles + halo and 16-byte alignment) — an estimate from the ROI table, not a counter
        R = 2
        box_bytes = ((((256 + 2 * R - 1 + 15) // 4 + 1) + 3) // 4 * 4) * 4 * (32 + 2 * R)
        tiles = np.ceil(rr["roi"][:, 3] / 32.0) * np.ceil(rr["roi"][:, 2] / 256.0)
        e2e = {"value": S / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(tiles.sum() * box_bytes), "h2d_bytes_note": "estimated: TMA boxes of the ROI tiles read in place from pinned host memory",
               "roi_bytes_per_step": int((rr["roi"][:, 2] * rr["roi"][:, 3]).sum()), "whole_image_bytes_per_step": S * W * H,
               "d2h_bytes_per_step": S * C.sizeof(mpe.MpeResult), "streams_updated_last_step": int(rr["updated"].sum()),
               "api": "mpe_streams_step_device on the device alias of a pinned host ring (zero-copy ingest) + result records D2H every step"}
    line = {"metric": f"frames/sec ({W}x{H}, {args.leds} LEDs, tracking mode, device-resident streams)", "value": S / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 + f64", "data": "synthetic",
            "config": {"workload": f"{S} streams ({Sd} distinct trajectories replayed), one frame per stream per step, ROI search + NN correspondences + checkCorrespondences + optimisePose",
                       "streams_updated_last_step": int(res["updated"].sum()), "mean_roi_pixels": float(np.mean(res["roi"][:, 2] * res["roi"][:, 3])),
                       "reinitialised_last_step": int(np.sum((res["flags"] & 8) != 0))},
            "e2e": e2e, "gpu_launches": launches}
    if not args.no_cpu:
        # CPU oracle in tracking mode, one thread, on one of the trajectories
        import cv2
        cv2.setNumThreads(1)
        from oracle import pose_oracle
        n_cpu = 1200
        sc = synth.make_stream_scene(n_cpu, n_leds=args.leds, width=W, height=H, seed=args.seed + 5)     # one long continuous trajectory
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        # the trajectory is played forwards and backwards (continuous motion, monotonic time stamps) until ~10 s have passed
        order = list(range(n_cpu)) + list(range(n_cpu - 2, 0, -1))
        tcur = 0.0
        for i in range(8):
            est.estimate_body_pose(sc.frames[order[i]], tcur); tcur += 1 / 60.0
        t0c = time.perf_counter()
        n, i = 0, 8
        while time.perf_counter() - t0c < 10.0:
            est.estimate_body_pose(sc.frames[order[i % len(order)]], tcur); tcur += 1 / 60.0
            i += 1; n += 1
        dtc = time.perf_counter() - t0c
        line["cpu_baseline"] = {"value": n / dtc, "unit": "frames/s", "cores": 1, "kind": "port",
                                "sample": f"{n} tracking-mode frames in {dtc:.1f} s, one thread (cv2 4.13 findLeds on the ROI + C++ oracle)"}
    print(json.dumps(line), flush=True)
    ctx.close()


# ------------------------------------------------------------------------------------------------ single-camera latency (secondary)
def run_latency(args):
    """What ONE camera sees (the way MPENode drives the reference: one image per callback, monocular_pose_estimator.cpp:133-159):
    per-image wall-clock latency of mpe_streams_step(n_streams=1) — H2D of the whole image, the tracking step replayed as one CUDA
    graph, D2H of the result record, synchronise — against the CPU oracle's estimateBodyPose on the same sequence, same host."""
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth
    from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
    from oracle import pose_oracle
    import cv2
    cv2.setNumThreads(1)
    W, H = args.width, args.height
    T = 40 + args.steps * 20
    sc = synth.make_stream_scene(T, n_leds=args.leds, width=W, height=H, seed=args.seed)
    out = {}
    # the camera driver's receive buffer: ONE slot, pinned or pageable, rewritten for every image (a stable address keeps the CUDA
    # graph of the step valid; pinned + AUTO ingest lets the kernels read the ROI in place once the stream is tracking)
    slots = {"pinned": torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy(), "pageable": np.empty((H, W), np.uint8)}
    for label, slot in slots.items():
        for graphs in (True, False):
            ctx = mpe.Context(0, 1, W, H)
            ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
            ctx.set_graph_replay(graphs)
            ctx.streams_reset(1)
            res = (mpe.MpeResult * 1)()
            tarr = np.zeros(1)
            tp = tarr.ctypes.data_as(C.POINTER(C.c_double))
            L, h = ctx.L, ctx.h
            ptr = C.c_void_p(slot.ctypes.data)
            lat, upd = [], 0
            l0 = ctx.launch_count()
            for t in range(T):
                tarr[0] = sc.times[t]
                slot[:] = sc.frames[t]               # the image arrives (not timed: the reference receives it the same way)
                t0 = time.perf_counter()
                rc = L.mpe_streams_step(h, ptr, W, W * H, W, H, 1, tp, res)
                t1 = time.perf_counter()
                assert rc == 0
                upd += res[0].updated
                if t >= 40:
                    lat.append((t1 - t0) * 1e6)
            lat = np.array(lat)
            st = ctx.ingest_stats()
            out[f"{label}_{'graph' if graphs else 'launches'}"] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                                                                    "mean_us": float(lat.mean()), "frames_updated": upd, "frames": T,
                                                                    "gpu_launches_per_frame": (ctx.launch_count() - l0) / T,
                                                                    "zero_copy_steps": st["zero_copy_steps"], "copy_steps": st["copy_steps"]}
            ctx.close()
    # cameras per call: n independent streams advanced by ONE mpe_streams_step (host images in, host records out)
    sweep = []
    for n in (1, 2, 4, 8, 16, 64, 256):
        ctx = mpe.Context(0, n, W, H)
        ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
        ctx.streams_reset(n)
        Tn = min(T, 120)
        hb = torch.empty((n, H, W), dtype=torch.uint8).pin_memory().numpy()
        res = (mpe.MpeResult * n)()
        tarr = np.zeros(n)
        tp = tarr.ctypes.data_as(C.POINTER(C.c_double))
        ptr = C.c_void_p(hb.ctypes.data)
        lat = []
        for t in range(Tn):
            hb[:] = sc.frames[t]                    # every camera sees the same sequence (results are per stream anyway)
            tarr[:] = sc.times[t]
            t0 = time.perf_counter()
            rc = ctx.L.mpe_streams_step(ctx.h, ptr, W, W * H, W, H, n, tp, res)
            t1 = time.perf_counter()
            assert rc == 0
            if t >= 20:
                lat.append((t1 - t0) * 1e6)
        assert all(res[i].updated for i in range(n))
        p50 = float(np.percentile(lat, 50))
        sweep.append({"cameras_per_call": n, "p50_us_per_call": p50, "us_per_image": p50 / n, "images_per_s": n / (p50 * 1e-6)})
        ctx.close()
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    lat = []
    for t in range(T):
        t0 = time.perf_counter()
        est.estimate_body_pose(sc.frames[t], sc.times[t])
        t1 = time.perf_counter()
        if t >= 40:
            lat.append((t1 - t0) * 1e6)
    lat = np.array(lat)
    cpu = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "mean_us": float(lat.mean()), "cores": 1, "kind": "port"}
    best = out["pinned_graph"]
    line = {"metric": f"single-camera latency per image ({W}x{H}, {args.leds} LEDs, tracking mode)", "value": best["p50_us"], "unit": "us",
            "n_gpus": 1, "higher_is_better": False, "data": "synthetic", "variants": out, "cameras_per_call_sweep": sweep, "cpu_baseline": cpu,
            "config": {"workload": f"one stream, {T} consecutive frames, one mpe_streams_step call per image (whole image H2D + graph replay + result D2H + sync)"}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="frames per GPU per step")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--leds", type=int, default=N_LEDS)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--mode", default="cold", choices=["cold", "tracking", "latency"],
                    help="cold (headline): every frame runs the full pipeline; tracking: device-resident streams with ROI search")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--contexts", type=int, default=2, help="batches in flight in the device-resident measurement (one context + stream each)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.mode == "tracking":
        run_tracking(args)
        return
    if args.mode == "latency":
        run_latency(args)
        return

    import torch
    import torch.distributed as dist
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth
    from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank)
    sampler.start()                      # started early: nvidia-smi needs ~1 s before its first sample
    W, H, B = args.width, args.height, args.batch
    scene = synth.make_cold_scene(B, n_leds=args.leds, width=W, height=H, seed=args.seed + 100000 * rank)
    host_frames = torch.from_numpy(scene.frames).pin_memory()          # B x H x W u8, pinned
    dev_frames = host_frames.to(dev, non_blocking=False)
    ctx = mpe.Context(local_rank, B, W, H)
    ctx.set_camera(scene.K, scene.D)
    ctx.set_params(scene.params)
    ctx.set_markers(scene.markers)
    stream = torch.cuda.Stream(device=dev)          # explicit stream: events and kernels share it (handle 0 would mean "own stream")
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # Two batches in flight: a second context on a second stream takes every other step, so that the low-occupancy tail of one
    # batch (thread-per-frame check / Gauss-Newton: < 2 warps per SM) overlaps the HBM-bound scan of the next (measured: 2.65 ->
    # 2.43 ms per 8192-frame batch).  One context per in-flight batch is the public-API way to do that; --contexts 1 switches it off.
    ctxs, streams = [ctx], [stream]
    for _ in range(1, max(1, args.contexts)):
        c2 = mpe.Context(local_rank, B, W, H)
        c2.set_camera(scene.K, scene.D); c2.set_params(scene.params); c2.set_markers(scene.markers)
        s2 = torch.cuda.Stream(device=dev)
        c2.set_stream(s2.cuda_stream)
        ctxs.append(c2); streams.append(s2)
    n_ctx = len(ctxs)
    rec_bytes = C.sizeof(mpe.MpeResult)
    # pose gather: the poses of this rank's frames, all-gathered over NCCL once per step.  It runs on a side stream,
    # double buffered, so that the collective of step i overlaps the kernels of step i+1 (nothing in the path waits for it).
    gather_in = [torch.zeros(B * 16, dtype=torch.float64, device=dev) for _ in range(2)]
    gather_out = [torch.zeros(world * B * 16, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    gstream = torch.cuda.Stream(device=dev) if world > 1 else None
    copied = [torch.cuda.Event() for _ in range(2)]
    gathered = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]

    def step_device(single=False):
        k = 0 if single else step_no[0] % n_ctx       # which context / stream takes this batch
        cx, sx = ctxs[k], streams[k]
        i = step_no[0] & 1
        step_no[0] += 1
        cx.estimate_batch_device_async(dev_frames.data_ptr(), W, W * H, W, H, B)
        if world > 1:
            if step_no[0] > 2:
                sx.wait_event(gathered[i])                # the gather that last read this buffer has finished
            cx.copy_poses_device(gather_in[i].data_ptr(), B)
            copied[i].record(sx)
            with torch.cuda.stream(gstream):
                gstream.wait_event(copied[i])
                dist.all_gather_into_tensor(gather_out[i], gather_in[i])
                gathered[i].record(gstream)

    def join():
        """everything enqueued so far, on either context's stream or the gather stream, is ordered before what `stream` gets next"""
        for sx in streams[1:]:
            stream.wait_stream(sx)
        if world > 1:
            stream.wait_stream(gstream)

    def barrier():
        join()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness spot check against the oracle (outside any timed region)
    res = results_to_arrays(ctx.estimate_batch_device(dev_frames.data_ptr(), W, W * H, W, H, B))
    n_updated = int(res["updated"].sum())
    if rank == 0:
        from oracle import pose_oracle
        for f in range(0, min(B, 8)):
            est = pose_oracle.PoseEstimatorOracle(scene.K, scene.D, scene.markers, scene.params)
            upd = est.estimate_body_pose(scene.frames[f], 0.0)
            assert bool(res[f]["updated"]) == upd, "GPU/oracle disagree on pose_updated"
            if upd:
                assert np.abs(res[f]["pose"].reshape(4, 4) - est.predicted_pose()).max() < 1e-6, "GPU/oracle pose mismatch"

    t_load0 = time.time()                # clocks are summarised over warm-up + timed steps (both under the same load)
    for _ in range(args.warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = sum(c.launch_count() for c in ctxs)
    e0.record(stream)
    for sx in streams[1:]:
        sx.wait_event(e0)                                 # the other stream starts inside the timed region too
    for _ in range(args.steps):
        step_device()
    join()                                                # the timed region ends when every batch and the last pose gather have landed
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches_timed = sum(c.launch_count() for c in ctxs) - launches1
    # per-kernel durations: a separate pass on ONE context with CUDA events around every stage; same load, so inside the clock window
    ctx.enable_kernel_timing(True)
    for _ in range(3):
        step_device(single=True)
    barrier()
    kt = ctx.kernel_times_ms()                      # last step's per-kernel CUDA-event durations (whole batch, stages back to back)
    ctx.enable_kernel_timing(False)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- e2e through the public host-buffer API (H2D + D2H inside)
    e2e = None
    if not args.no_e2e:
        hp = host_frames.numpy()
        for _ in range(2):
            ctx.estimate_batch(hp)
        barrier()
        e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        steps_e2e = max(3, min(args.steps, 10))
        e2[0].record(stream)
        t0 = time.perf_counter()
        for _ in range(steps_e2e):
            r = ctx.estimate_batch(hp)
        e2[1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms_e2e = max(e2[0].elapsed_time(e2[1]), wall * 1e3)
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item()) / steps_e2e
        e2e = {"value": world * B / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
               "d2h_bytes_per_step": B * rec_bytes, "ms_per_step": ms_e2e, "steps": steps_e2e,
               "h2d_gbs": B * W * H / (ms_e2e * 1e-3) / 1e9,      # PCIe-bound: compare with the box's pinned H2D copy rate (55.6 GB/s measured, tests/probes/pcie_probe.py)
               "api": "mpe_estimate_batch (pinned host frames -> host mpe_result records)"}
        assert int(results_to_arrays(r)["updated"].sum()) == n_updated

    if rank == 0:
        peak, peak_src = load_peaks()
        names = ["scan (K1a)", "extract_blobs (K1b)", "p3p_sweep (K2)", "check+refine (K3)", "blur_tiles (K1c)"]
        ksum = sum(kt)
        alg_bytes = B * W * H                                        # ROI bytes read once
        k1_gbs = alg_bytes / (kt[0] * 1e-3) / 1e9 if kt[0] > 0 else 0.0
        kernels = []
        for nme, ms in zip(names, kt):
            kernels.append({"name": nme, "ms": ms, "share_of_kernel_time": ms / ksum if ksum > 0 else None})
        kernels[0].update({"bound": "hbm", "achieved_gbs": k1_gbs, "frac": k1_gbs / peak})
        kernels[1].update({"bound": "latency (sparse contour tracing)"})
        kernels[2].update({"bound": "fp64 alu", "p3p_solves_per_s": B * 600 / (kt[2] * 1e-3) if kt[2] > 0 and args.leds == 5 else None})
        kernels[3].update({"bound": "fp64 alu / latency (dependent GN chain)"})
        kernels[4].update({"bound": "latency (sparse exact blur of hot tiles)"})
        line = {
            "metric": "frames/sec (752x480, 5 LEDs, cold full pipeline)" if (W, H, args.leds) == (752, 480, 5) else f"frames/sec ({W}x{H}, {args.leds} LEDs, cold)",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (findLeds, fixed point) + f64 (P3P, Gauss-Newton)",
            "data": "synthetic",
            "config": {"workload": f"{W}x{H} synthetic stream, {args.leds} LEDs, cold mode: whole-image findLeds + initialise + checkCorrespondences + optimisePose for every frame",
                       "frames_per_gpu_per_step": B, "global_frames_per_step": world * B, "parallelism": f"frames sharded over {world} GPU(s), NCCL pose all-gather per step" if world > 1 else "1 GPU",
                       "l2": f"batch of {B * W * H / 1e6:.0f} MB per GPU > 126 MB L2 (inputs larger than L2, no flush needed)",
                       "batches_in_flight": n_ctx,
                       "frames_with_pose": n_updated},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_timed,
            "roofline": {"kernel": "scan_kernel (K1a: TMA-streamed threshold scan of every ROI byte; findLeds hot loop)", "bound": "hbm", "achieved": k1_gbs, "peak": peak, "unit": "GB/s",
                         "frac": k1_gbs / peak, "peak_source": peak_src, "traffic": load_traffic(B, W, H),
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kt[0]},
            "dominant_kernel": names[int(np.argmax(kt))],
            "kernel_times_note": "stage times from a separate pass on one context with CUDA events around every stage (sum = %.3f ms); in the timed steps "
                                 "%d batches are in flight on %d streams, so ms_per_step can be smaller than that sum" % (ksum, n_ctx, n_ctx),
            "kernels": kernels,
        }
        if not args.no_cpu and world >= 1:
            fps, n, dt = cpu_single_thread(scene)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} frames of the same batch in {dt:.1f} s, one thread (cv2 4.13 findLeds + C++ oracle of the pose path)",
                                    "host_cpus": os.cpu_count(), "host_cpus_usable": usable_cores()}
        print(json.dumps(line), flush=True)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
