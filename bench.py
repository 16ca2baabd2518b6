#!/usr/bin/env python
"""bench.py — frames/sec of the per-frame hot path (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[1]): 752x480 synthetic stream, 5 LEDs, full pipeline in COLD mode — every frame
runs whole-image findLeds + initialise (600 P3P problems, 2400 hypotheses) + checkCorrespondences + optimisePose.
One "step" = one pass of that path over `--batches-per-step` batches of `--batch` frames per GPU (default 16 x 8192 frames,
so that the timed region of the default 20 steps lasts about half a second).

  value     frames/s with the frames already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same metric through the C-ABI call a user makes (mpe_estimate_batch) with HOST (pinned) frames:
            H2D of every frame and D2H of every result record inside the timed region
  roofline  the HBM-bound kernel (scan_kernel of findLeds): algorithmic bytes = W*H per frame / its CUDA-event duration,
            against MEASURED_PEAKS.json; `roofline_fp64` is the same for the FP64-bound sweep (K2): exact FP64 operation
            count of the CPU restatement (oracle built on a counting double) / its duration, against a measured DFMA peak;
            `kernels` lists every kernel's time and share
  cpu_baseline  the CPU oracle (cv2 findLeds + C++ pose restatement), one thread, bounded sample, same frames
  extra     (1 GPU only) the other configurations of BASELINE.json measured in the same run, each with its own clock window and
            CPU baseline: tracking mode (device-resident streams, zero-copy e2e), one camera's per-image latency, 1920x1080,
            8 LEDs

`--impl reference` times the reference's CPU path on all host cores: the faster of the oracle port and of the UNMODIFIED
reference sources built in oracle/_ref (stand-in Eigen, OpenCV calls bridged to cv2) — the faster one is the conservative
denominator.  Multi-GPU: one process per GPU (torchrun), frames sharded, no data-path collective; the mpe_result records are
all-gathered over NCCL once per batch on a side stream and verified after the timed region.
`--mode tracking --gpus N` is BASELINE config 5 (streams sharded over the GPUs, NCCL gather of the records).
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
METRIC = "frames/sec (752x480, 5 LEDs, cold full pipeline)"


# ------------------------------------------------------------------------------------------------ shared bits
def make_config(args):
    """The workload description — identical, key for key, in the GPU arm and in the reference arm."""
    return {"workload": f"{args.width}x{args.height} synthetic stream, {args.leds} LEDs, cold mode: whole-image findLeds + initialise + "
                        "checkCorrespondences + optimisePose for every frame",
            "width": args.width, "height": args.height, "leds": args.leds, "mode": "cold",
            "frames_per_gpu_per_step": args.batch * args.batches_per_step, "batch": args.batch, "batches_per_step": args.batches_per_step,
            "distinct_frames_per_gpu": args.batch, "seed": args.seed,
            "l2": f"every batch reads {args.batch * args.width * args.height / 1e6:.0f} MB of distinct frames per GPU (> 126 MB L2): inputs larger than L2, no flush"}


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
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy: the kernel is timed alone between events)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, n_gpus=1, enabled=True):
        """gpu_index: first GPU to watch, n_gpus: how many (one nvidia-smi process for all of them: with one poller per rank,
        eight of them queried the driver 50 times a second each while the kernels were being launched)."""
        self.gpus = list(range(gpu_index, gpu_index + n_gpus))
        self.enabled = enabled
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        if not self.enabled:
            return
        try:
            period = "20" if len(self.gpus) == 1 else "50"
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", period,
                                       "-i", ",".join(str(g) for g in self.gpus)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def window(self, t_begin, t_end):
        """Summary of the samples taken between two wall-clock times (the sampler keeps running)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            out["error"] = "nvidia-smi not available"
            return out
        self.f.flush()
        per = {}                                   # GPU index -> (sm samples, max clocks, power samples)
        reasons = set()
        per_reasons = {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.f.name) as fh:
            for line in fh.read().splitlines():
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_begin - 0.02 or ts > t_end + 0.02:
                        continue
                    idx = int(parts[1])
                    sm_v, mx_v = float(parts[2]), float(parts[3])
                    pw_v = float(parts[4])
                except ValueError:
                    continue
                sm, mx, pw = per.setdefault(idx, ([], [], []))
                sm.append(sm_v); mx.append(mx_v); pw.append(pw_v)
                for n, v in zip(names, parts[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(n)                         # plain names in `reasons` (what the driver looks for) ...
                        per_reasons.setdefault(idx, set()).add(n)   # ... which GPU it was goes into per_gpu
        if not per:
            out["error"] = "no nvidia-smi samples in the window"
        else:
            med = {i: statistics.median(v[0]) for i, v in per.items()}
            out.update({"sm_mhz": min(med.values()), "sm_max_mhz": max(max(v[1]) for v in per.values()), "samples": sum(len(v[0]) for v in per.values()),
                        "power_w_max": max(max(v[2]) for v in per.values()), "window_s": round(t_end - t_begin, 3)})
            if len(self.gpus) > 1:                 # sm_mhz above is the slowest GPU's median
                out["per_gpu"] = [{"gpu": i, "sm_mhz": med[i], "power_w_max": max(per[i][2]), "reasons": sorted(per_reasons.get(i, ()))} for i in sorted(per)]
        out["reasons"] = sorted(reasons)
        return out

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass


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


def bind_near_gpu(local_rank):
    """Pin this process (and therefore the first touch of its page-locked frame buffer) to the CPUs next to its GPU, so that at
    N > 1 the ranks' H2D streams do not all cross the same socket link.  Returns a description for the JSON line."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            info = {"bound": True, "cpus": len(target), "first_cpu": min(target), "last_cpu": max(target)}
        try:
            info["numa_node"] = int(open(f"/sys/bus/pci/devices/{pynvml.nvmlDeviceGetPciInfo(h).busId.decode().lower()[4:]}/numa_node").read())
        except Exception:
            pass
    except Exception as e:                                   # no NVML / not permitted: keep the default placement
        info["error"] = str(e)[:80]
    return info


# ------------------------------------------------------------------------------------------------ CPU arms
def _cpu_worker_init(K, D, markers, params, kind):
    import cv2
    cv2.setNumThreads(1)
    global _W
    from oracle import pose_oracle
    cls = pose_oracle.PoseEstimatorOracle
    if kind == "reference":
        from oracle import ref_pose
        cls = ref_pose.PoseEstimatorRef
    _W = dict(K=K, D=D, markers=markers, params=params, cls=cls)


_FRAMES = None   # set in the parent before the fork so that workers get the frames without any IPC


def _cpu_worker_run(idx):
    n_upd = 0
    for i in idx:
        fr = _FRAMES[i % len(_FRAMES)]
        est = _W["cls"](_W["K"], _W["D"], _W["markers"], _W["params"])
        n_upd += int(est.estimate_body_pose(fr, 0.0))
    return n_upd


def cpu_single_thread(scene, max_seconds=12.0, max_frames=20000, cls=None):
    import cv2
    cv2.setNumThreads(1)
    from oracle import pose_oracle
    cls = cls or pose_oracle.PoseEstimatorOracle
    n = 0
    for f in range(min(5, len(scene.frames))):
        cls(scene.K, scene.D, scene.markers, scene.params).estimate_body_pose(scene.frames[f], 0.0)
    t0 = time.perf_counter()
    while n < max_frames and time.perf_counter() - t0 < max_seconds:
        est = cls(scene.K, scene.D, scene.markers, scene.params)
        est.estimate_body_pose(scene.frames[n % len(scene.frames)], 0.0)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference_arm(args):
    """The reference's CPU implementation of the path on all host cores; rank 0 only.  Two builds exist here: the oracle port
    (cv2 4.13 findLeds + C++ restatement) and, when oracle/_ref was built, the UNMODIFIED reference sources (stand-in Eigen,
    OpenCV calls bridged to cv2).  Both are timed; the line's value is the FASTER one (the conservative denominator)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from rpg_monocular_pose_estimator_b200 import synth
    from oracle import ref_pose
    cores = usable_cores()
    per_core = 48
    sample = cores * per_core
    scene = synth.make_cold_scene(min(sample, 512), n_leds=args.leds, width=args.width, height=args.height, seed=args.seed)
    global _FRAMES
    _FRAMES = scene.frames
    chunks = [list(range(i, sample, cores)) for i in range(cores)]
    ctx = mp.get_context("fork")
    kinds = ["port"] + (["reference"] if ref_pose.available() else [])
    runs = {}
    for kind in kinds:
        with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(scene.K, scene.D, scene.markers, scene.params, kind)) as pool:
            for _ in range(max(args.warmup, 1)):
                pool.map(_cpu_worker_run, chunks)
            t0 = time.perf_counter()
            upd = 0
            for _ in range(args.steps):
                upd += sum(pool.map(_cpu_worker_run, chunks))
            dt = time.perf_counter() - t0
        runs[kind] = {"fps": args.steps * sample / dt, "dt": dt, "updated_per_step": upd // args.steps}
    best = max(runs, key=lambda k: runs[k]["fps"])
    fps, dt = runs[best]["fps"], runs[best]["dt"]
    what = {"port": "cv2 4.13 findLeds + C++ restatement of the pose path (oracle/pose_oracle.cpp)",
            "reference": "unmodified reference sources built in oracle/_ref (stand-in Eigen; OpenCV calls bridged to cv2 4.13)"}
    line = {
        "impl": "reference", "metric": METRIC if (args.width, args.height, args.leds) == (WIDTH, HEIGHT, N_LEDS) else f"frames/sec ({args.width}x{args.height}, {args.leds} LEDs, cold)",
        "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (findLeds, fixed point) + f64 (P3P, Gauss-Newton)", "data": "synthetic",
        "config": make_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": best,
                         "sample": f"{sample} frames per step x {args.steps} steps on {cores} processes (one OpenCV thread each): {what[best]}",
                         "all_builds_fps": {k: v["fps"] for k, v in runs.items()}, "frames_updated_per_step": runs[best]["updated_per_step"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ tracking mode (config 2b / config 5)
def tracking_measure(args, dev_index, S, Sd, steps, warmup, sampler, dist=None, world=1, rank=0, cpu=True, e2e=True, verify_streams=0):
    """Steady-state tracking throughput: S streams per GPU, each an independent PoseEstimator whose state lives on the GPU; one step =
    one frame for every stream (mpe_streams_step_device).  Sd distinct synthetic trajectories of T frames are replayed by the S
    streams through the frame map.  With world > 1 the streams are sharded over the ranks and the mpe_result records are gathered
    over NCCL every step (BASELINE config 5); `verify_streams` > 0: rank 0 also advances the first streams of EVERY rank itself and
    compares them with the gathered table."""
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth, sharding
    from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
    dev = torch.device("cuda", dev_index)
    W, H = args.width, args.height
    T = 24                                    # frames per recorded trajectory; replayed forwards and backwards (continuous motion, monotonic
    rec = C.sizeof(mpe.MpeResult)             # time stamps), so that the number of timed steps does not depend on the recording length
    n_sched = 8 + warmup + steps + 1
    pingpong = list(range(T)) + list(range(T - 2, 0, -1))
    def frame_of(t):
        return pingpong[t % len(pingpong)]
    def trajectories(r, n=Sd):
        return [synth.make_stream_scene(T, n_leds=args.leds, width=W, height=H, seed=args.seed + 17 * s + 7919 * r) for s in range(n)]
    seqs = trajectories(rank)
    buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).to(dev)          # (Sd*T) x H x W
    ctx = mpe.Context(dev_index, S, W, H)
    ctx.set_camera(seqs[0].K, seqs[0].D); ctx.set_params(seqs[0].params); ctx.set_markers(seqs[0].markers)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    ctx.streams_reset(S)
    base = (torch.arange(S, dtype=torch.int32, device=dev) % Sd) * T
    fmap = base.clone()                       # one buffer, advanced in place: a stable key lets the step replay as a CUDA graph
    times = [np.full(S, t / 60.0) for t in range(n_sched + 1)]
    ctx.streams_set_frame_map(fmap.data_ptr(), Sd * T)
    g_in = torch.zeros((S, rec), dtype=torch.uint8, device=dev)
    g_out = torch.zeros((world * S, rec), dtype=torch.uint8, device=dev) if world > 1 else None
    gev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    gather_ms = []

    def step(t, frames_ptr, fetch=False, gather=False):
        fmap.copy_(base + frame_of(t))
        r = ctx.streams_step_device(frames_ptr, W, W * H, W, H, times[t], fetch=fetch)
        if gather and world > 1:
            ctx.copy_results_device(g_in.data_ptr(), S)
            gev[0].record(stream)
            sharding.gather_records(g_in, world, dist, out=g_out)
            gev[1].record(stream)
        return r

    t = 0
    for _ in range(8 + warmup):               # cold start + settle into tracking (it_since_initialized_ == 2), untimed
        step(t, buf.data_ptr(), gather=True); t += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        step(t, buf.data_ptr(), gather=True); t += 1
        if world > 1:
            gev[1].synchronize(); gather_ms.append(gev[0].elapsed_time(gev[1]))
    e1.record(stream)
    torch.cuda.synchronize()
    tw1 = time.time()
    ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item()) / steps
    res = results_to_arrays(ctx.fetch_results(S))
    launches = ctx.launch_count() - l0
    out = {"metric": f"frames/sec ({W}x{H}, {args.leds} LEDs, tracking mode, device-resident streams)", "value": world * S / (ms * 1e-3), "unit": "frames/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8 (findLeds, fixed point) + f64", "data": "synthetic",
           "config": {"workload": f"{S} streams per GPU ({Sd} distinct trajectories replayed), one frame per stream per step: predictWithROI + ROI findLeds + "
                                  "findCorrespondences + checkCorrespondences + optimisePose (re-initialisation when the check fails)",
                      "width": W, "height": H, "leds": args.leds, "mode": "tracking", "streams_per_gpu": S, "distinct_trajectories": Sd, "seed": args.seed},
           "streams_updated_last_step": int(res["updated"].sum()), "mean_roi_pixels": float(np.mean(res["roi"][:, 2] * res["roi"][:, 3])),
           "reinitialised_last_step": int(np.sum((res["flags"] & 8) != 0)),
           "clocks": sampler.window(tw0, tw1), "gpu_launches": launches}
    if world > 1:
        ok = sharding.verify_gather(g_out, g_in, rank, world, dist)
        okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        out["gather"] = {"collective": "all_gather_into_tensor of the mpe_result records, every step, on the step's stream", "bytes_per_rank_per_step": S * rec,
                         "ms_per_step_rank0": float(np.mean(gather_ms)), "verified_all_ranks": bool(okt.item())}
        if verify_streams > 0:
            # the gathered table against ONE rank advancing the same streams by itself: rank 0 replays the first streams of every rank
            n_v = min(verify_streams, S, Sd)
            table = g_out.cpu().numpy().reshape(world, S, rec)[:, :n_v]
            t_last = t - 1
            if rank == 0:
                vctx = mpe.Context(dev_index, world * n_v, W, H)
                vctx.set_camera(seqs[0].K, seqs[0].D); vctx.set_params(seqs[0].params); vctx.set_markers(seqs[0].markers)
                vs = torch.cuda.Stream(device=dev); vctx.set_stream(vs.cuda_stream)
                allseq = [trajectories(r, n_v) for r in range(world)]
                vbuf = torch.from_numpy(np.stack([sc.frames[frame_of(tt)] for tt in range(t_last + 1) for tr in allseq for sc in tr])).to(dev)   # t x (world*n_v) x H x W
                vctx.streams_reset(world * n_v)
                vctx.streams_set_frame_map(0, 0)
                with torch.cuda.stream(vs):
                    for tt in range(t_last + 1):
                        rr = vctx.streams_step_device(vbuf[tt * world * n_v].data_ptr(), W, W * H, W, H, np.full(world * n_v, tt / 60.0), fetch=(tt == t_last))
                one = np.frombuffer(rr, dtype=np.uint8).reshape(world, n_v, rec)
                out["gather"]["table_equals_single_rank_run"] = bool(np.array_equal(one, table))
                out["gather"]["table_check_streams"] = world * n_v
                vctx.close()
    # ---- e2e: the same streams fed from a PINNED HOST ring (zero-copy ingest): the findLeds kernels read their ROI tiles in place
    # over PCIe, every step ends with the D2H copy of all result records and a synchronise (wall clock and CUDA events, the larger)
    if e2e:
        host_buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).pin_memory()
        ctx.streams_reset(S)
        t = 0
        for _ in range(8 + warmup):
            step(t, host_buf.data_ptr(), fetch=True); t += 1
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        w0 = time.perf_counter()
        for _ in range(steps):
            r = step(t, host_buf.data_ptr(), fetch=True); t += 1
        h1.record(stream)
        torch.cuda.synchronize()
        ms_e2e = max((time.perf_counter() - w0) * 1e3, h0.elapsed_time(h1)) / steps
        mt = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        ms_e2e = float(mt.item())
        rr = results_to_arrays(r)
        # bytes that cross PCIe per step: whole TMA boxes of the tiles each ROI touches (32-row strips + 2R halo rows, 256-px column
        # tiles + halo and 16-byte alignment) — an estimate from the ROI table, not a counter
        R = 2
        box_bytes = ((((256 + 2 * R - 1 + 15) // 4 + 1) + 3) // 4 * 4) * 4 * (32 + 2 * R)
        tiles = np.ceil(rr["roi"][:, 3] / 32.0) * np.ceil(rr["roi"][:, 2] / 256.0)
        out["e2e"] = {"value": world * S / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": int(tiles.sum() * box_bytes), "h2d_bytes_note": "estimated: TMA boxes of the ROI tiles read in place from pinned host memory",
                      "roi_bytes_per_step": int((rr["roi"][:, 2] * rr["roi"][:, 3]).sum()), "whole_image_bytes_per_step": S * W * H,
                      "d2h_bytes_per_step": S * rec, "streams_updated_last_step": int(rr["updated"].sum()),
                      "api": "mpe_streams_step_device on the device alias of a pinned host ring (zero-copy ingest) + result records D2H every step"}
    if cpu and rank == 0:
        import cv2
        cv2.setNumThreads(1)
        from oracle import pose_oracle
        n_cpu = 600
        sc = synth.make_stream_scene(n_cpu, n_leds=args.leds, width=W, height=H, seed=args.seed + 5)     # one long continuous trajectory
        est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
        order = list(range(n_cpu)) + list(range(n_cpu - 2, 0, -1))      # forwards and backwards: continuous motion, monotonic time stamps
        tcur = 0.0
        for i in range(8):
            est.estimate_body_pose(sc.frames[order[i]], tcur); tcur += 1 / 60.0
        t0c = time.perf_counter()
        n, i = 0, 8
        while time.perf_counter() - t0c < args.cpu_seconds * 0.6:
            est.estimate_body_pose(sc.frames[order[i % len(order)]], tcur); tcur += 1 / 60.0
            i += 1; n += 1
        dtc = time.perf_counter() - t0c
        out["cpu_baseline"] = {"value": n / dtc, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": f"{n} tracking-mode frames in {dtc:.1f} s, one thread (cv2 4.13 findLeds on the ROI + C++ oracle)"}
    ctx.close()
    del buf
    torch.cuda.empty_cache()
    return out


def run_tracking(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(0, n_gpus=world, enabled=(local_rank == 0)) if world > 1 else ClockSampler(local_rank)
    sampler.start(); time.sleep(1.2)
    line = tracking_measure(args, local_rank, args.batch, 512, args.steps, args.warmup, sampler, dist if world > 1 else None, world, rank,
                            cpu=not args.no_cpu, e2e=not args.no_e2e, verify_streams=32 if world > 1 else 0)
    sampler.stop()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ single-camera latency
def latency_measure(args, sampler, n_frames=240, sweep=(1, 2, 8, 64), variants=True):
    """What ONE camera sees (the way MPENode drives the reference: one image per callback, monocular_pose_estimator.cpp:133-159):
    per-image wall-clock latency of mpe_streams_step(n_streams=1) — H2D of the image (or in-place ROI reads), the tracking step
    replayed as one CUDA graph, D2H of the result record, synchronise — against the CPU oracle's estimateBodyPose, same host."""
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth
    from oracle import pose_oracle
    import cv2
    cv2.setNumThreads(1)
    W, H = args.width, args.height
    T = 40 + n_frames
    sc = synth.make_stream_scene(T, n_leds=args.leds, width=W, height=H, seed=args.seed)
    out = {}
    tw0 = time.time()
    slots = {"pinned": torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy()}
    if variants:
        slots["pageable"] = np.empty((H, W), np.uint8)
    for label, slot in slots.items():
        for graphs in ((True, False) if variants else (True,)):
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
    cams = []
    for n in sweep:                              # cameras per call: n independent streams advanced by ONE mpe_streams_step
        ctx = mpe.Context(0, n, W, H)
        ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
        ctx.streams_reset(n)
        Tn = min(T, 100)
        hb = torch.empty((n, H, W), dtype=torch.uint8).pin_memory().numpy()
        res = (mpe.MpeResult * n)()
        tarr = np.zeros(n)
        tp = tarr.ctypes.data_as(C.POINTER(C.c_double))
        ptr = C.c_void_p(hb.ctypes.data)
        lat = []
        for t in range(Tn):
            hb[:] = sc.frames[t]
            tarr[:] = sc.times[t]
            t0 = time.perf_counter()
            rc = ctx.L.mpe_streams_step(ctx.h, ptr, W, W * H, W, H, n, tp, res)
            t1 = time.perf_counter()
            assert rc == 0
            if t >= 20:
                lat.append((t1 - t0) * 1e6)
        assert all(res[i].updated for i in range(n))
        p50 = float(np.percentile(lat, 50))
        cams.append({"cameras_per_call": n, "p50_us_per_call": p50, "us_per_image": p50 / n, "images_per_s": n / (p50 * 1e-6)})
        ctx.close()
    tw1 = time.time()
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    lat = []
    for t in range(T):
        t0 = time.perf_counter()
        est.estimate_body_pose(sc.frames[t], sc.times[t])
        t1 = time.perf_counter()
        if t >= 40:
            lat.append((t1 - t0) * 1e6)
    lat = np.array(lat)
    cpu = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "mean_us": float(lat.mean()), "cores": 1, "kind": "port",
           "value": float(np.percentile(lat, 50)), "unit": "us", "sample": f"{len(lat)} consecutive tracking-mode frames, one thread"}
    best = out["pinned_graph"]
    return {"metric": f"single-camera latency per image ({W}x{H}, {args.leds} LEDs, tracking mode)", "value": best["p50_us"], "unit": "us",
            "n_gpus": 1, "higher_is_better": False, "data": "synthetic", "variants": out, "cameras_per_call_sweep": cams, "cpu_baseline": cpu,
            "clocks": sampler.window(tw0, tw1),
            "config": {"workload": f"one stream, {T} consecutive frames, one mpe_streams_step call per image (image ingest + graph replay + result D2H + sync)",
                       "width": W, "height": H, "leds": args.leds, "mode": "latency"}}


def run_latency(args):
    import torch
    torch.cuda.set_device(0)
    sampler = ClockSampler(0); sampler.start(); time.sleep(1.2)
    line = latency_measure(args, sampler, n_frames=40 + args.steps * 20, sweep=(1, 2, 4, 8, 16, 64, 256))
    sampler.stop()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ cold mode (headline)
def k2_op_counts(scene, n_frames=12):
    """Exact FP64 operation counts of initialise() and optimisePose() per frame: the CPU restatement compiled on a counting double
    (oracle/counted_double.h), fed with the oracle's own detections of the first frames of the batch."""
    from oracle import pose_oracle, find_leds_cv2
    p = scene.params
    init, opt, special, n_ok = [], [], [], 0
    for f in range(min(n_frames, len(scene.frames))):
        px, _ = find_leds_cv2.find_leds(scene.frames[f], (0, 0, scene.width, scene.height), p.threshold_value, p.gaussian_sigma, p.min_blob_area,
                                        p.max_blob_area, p.max_width_height_distortion, p.max_circular_distortion, scene.K, scene.D)
        if px is None or len(px) < 4:
            continue
        r = pose_oracle.count_ops(scene.K, scene.D, scene.markers, p, px)
        init.append(r["initialise"]["flops"]); special.append(r["initialise"]["special"])
        if r["ok"]:
            opt.append(r["optimise"]["flops"]); n_ok += 1
    return {"initialise_flops_per_frame": float(np.mean(init)), "initialise_transcendental_calls_per_frame": float(np.mean(special)),
            "optimise_flops_per_frame": float(np.mean(opt)) if opt else None, "frames_counted": len(init),
            "what": "adds + multiplies + divides + square roots + comparisons on doubles, counted by oracle/libpose_oracle_counted.so "
                    "(the restatement compiled on a counting double); library transcendental calls listed separately"}


def cold_measure(args, local_rank, world, rank, dist, sampler, B, bps, steps, warmup, leds, W, H, seed, n_ctx_req, do_e2e, do_cpu, cpu_seconds, check_frames,
                 do_gather=True, bind_info=None):
    import torch
    import rpg_monocular_pose_estimator_b200 as mpe
    from rpg_monocular_pose_estimator_b200 import synth, sharding
    from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays
    dev = torch.device("cuda", local_rank)
    scene = synth.make_cold_scene(B, n_leds=leds, width=W, height=H, seed=seed + 100000 * rank)
    host_frames = torch.from_numpy(scene.frames).pin_memory()          # B x H x W u8, pinned (first touched on the CPUs next to this GPU)
    dev_frames = host_frames.to(dev, non_blocking=False)
    def new_ctx():
        c = mpe.Context(local_rank, B, W, H)
        c.set_camera(scene.K, scene.D); c.set_params(scene.params); c.set_markers(scene.markers)
        return c
    ctx = new_ctx()
    stream = torch.cuda.Stream(device=dev)          # explicit stream: events and kernels share it (handle 0 would mean "own stream")
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # Two batches in flight: a second context on a second stream takes every other batch, so that the low-occupancy tail of one
    # batch overlaps the HBM-bound scan of the next.  One context per in-flight batch is the public-API way to do that.
    ctxs, streams = [ctx], [stream]
    for _ in range(1, max(1, n_ctx_req)):
        c2 = new_ctx()
        s2 = torch.cuda.Stream(device=dev)
        c2.set_stream(s2.cuda_stream)
        ctxs.append(c2); streams.append(s2)
    n_ctx = len(ctxs)
    rec = C.sizeof(mpe.MpeResult)
    # record gather: the mpe_result records of this rank's batch, all-gathered over NCCL once per batch on a side stream,
    # double buffered, so that the collective of batch i overlaps the kernels of batch i+1 (nothing in the path waits for it).
    gather = do_gather and world > 1 and not getattr(args, "no_gather", False)
    n_slots = max(2, int(getattr(args, "gather_slots", 2)))   # ring of gather buffers: a batch waits only for the gather n_slots batches back
    g_in = [torch.zeros((B, rec), dtype=torch.uint8, device=dev) for _ in range(n_slots)] if gather else None
    g_out = [torch.zeros((world * B, rec), dtype=torch.uint8, device=dev) for _ in range(n_slots)] if gather else None
    gstream = torch.cuda.Stream(device=dev) if gather else None
    copied = [torch.cuda.Event() for _ in range(n_slots)]
    gathered = [torch.cuda.Event() for _ in range(n_slots)]
    g_t0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_slots)]
    g_t1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_slots)]
    batch_no = [0]
    gather_on = [gather]

    def batch_device(single=False):
        k = 0 if single else batch_no[0] % n_ctx       # which context / stream takes this batch
        cx, sx = ctxs[k], streams[k]
        i = batch_no[0] % n_slots
        batch_no[0] += 1
        cx.estimate_batch_device_async(dev_frames.data_ptr(), W, W * H, W, H, B)
        if gather_on[0]:
            if batch_no[0] > n_slots:
                sx.wait_event(gathered[i])                # the gather that last read this buffer has finished
            cx.copy_results_device(g_in[i].data_ptr(), B)
            copied[i].record(sx)
            with torch.cuda.stream(gstream):
                gstream.wait_event(copied[i])
                g_t0[i].record(gstream)
                sharding.gather_records(g_in[i], world, dist, out=g_out[i])
                g_t1[i].record(gstream)
                gathered[i].record(gstream)

    def join():
        """everything enqueued so far, on either context's stream or the gather stream, is ordered before what `stream` gets next"""
        for sx in streams[1:]:
            stream.wait_stream(sx)
        if gather:
            stream.wait_stream(gstream)

    def barrier():
        join()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness against the oracle (outside any timed region): update flag of `check_frames` frames, pose of those updated
    res = results_to_arrays(ctx.estimate_batch_device(dev_frames.data_ptr(), W, W * H, W, H, B))
    n_updated = int(res["updated"].sum())
    oracle_check = None
    if rank == 0 and check_frames > 0:
        from oracle import pose_oracle
        n_chk = min(B, check_frames)
        n_or = 0
        for f in range(n_chk):
            est = pose_oracle.PoseEstimatorOracle(scene.K, scene.D, scene.markers, scene.params)
            upd = est.estimate_body_pose(scene.frames[f], 0.0)
            n_or += int(upd)
            assert bool(res[f]["updated"]) == upd, f"GPU/oracle disagree on pose_updated (frame {f})"
            if upd:
                assert np.abs(res[f]["pose"].reshape(4, 4) - est.predicted_pose()).max() < 1e-6, f"GPU/oracle pose mismatch (frame {f})"
                assert int(res[f]["gn_iters"]) == est.gn_iterations(), f"GPU/oracle Gauss-Newton iteration count differs (frame {f})"
        assert int(res["updated"][:n_chk].sum()) == n_or
        oracle_check = {"frames": n_chk, "frames_with_pose_gpu": int(res["updated"][:n_chk].sum()), "frames_with_pose_oracle": n_or,
                        "checked": "pose_updated of every frame; pose within 1e-6 and equal Gauss-Newton iteration count where updated"}

    t_load0 = time.time()                # clocks are summarised over warm-up + timed steps (both under the same load)
    for _ in range(max(warmup, 3)):
        for _b in range(bps):
            batch_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = sum(c.launch_count() for c in ctxs)
    e0.record(stream)
    for sx in streams[1:]:
        sx.wait_event(e0)                                 # the other stream starts inside the timed region too
    th0 = time.perf_counter()
    for _ in range(steps):
        for _b in range(bps):
            batch_device()
    th1 = time.perf_counter()                             # host time to enqueue the timed batches (launches are asynchronous)
    join()                                                # the timed region ends when every batch and the last record gather have landed
    e1.record(stream)
    barrier()
    t_load1 = time.time()
    ms_total = e0.elapsed_time(e1)
    host_enqueue_ms_per_batch = (th1 - th0) * 1e3 / (steps * bps)
    launches_timed = sum(c.launch_count() for c in ctxs) - launches1
    gather_info = None
    if gather:
        gms = float(np.mean([g_t0[i].elapsed_time(g_t1[i]) for i in range(n_slots)]))
        oks = [sharding.verify_gather(g_out[i], g_in[i], rank, world, dist) for i in range(n_slots)]
        okt = torch.tensor([1 if all(oks) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_info = {"collective": "all_gather_into_tensor of the batch's mpe_result records (NCCL), side stream, ring of %d buffers, once per batch" % n_slots,
                       "bytes_per_rank_per_batch": B * rec, "ms_per_gather_rank0": gms, "verified_all_ranks": bool(okt.item()),
                       "verification": "own block byte for byte + checksum exchange for the other ranks' blocks (sharding.verify_gather), both buffers, after the timed region"}
    # per-kernel durations: a separate pass on ONE context with CUDA events around every stage, the record gather switched off
    # (an NCCL kernel resident on the SMs would push the persistent scan CTAs into a second wave and spoil the roofline figure)
    gather_on[0] = False
    ctx.enable_kernel_timing(True)
    for _ in range(3):
        batch_device(single=True)
    barrier()
    kt = ctx.kernel_times_ms()                      # last batch's per-kernel CUDA-event durations (stages back to back)
    ctx.enable_kernel_timing(False)
    fp64_peak = ctx.probe_fp64_peak() if rank == 0 else None
    clocks = sampler.window(t_load0, t_load1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    ms_step_per_rank = None
    if world > 1:
        allt = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, t)
        ms_step_per_rank = [float(x) / steps for x in allt.tolist()]   # the job's time is the slowest rank's
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    value = world * B * bps / (ms_step * 1e-3)

    # ---- e2e through the public host-buffer API (H2D + D2H inside)
    e2e = None
    if do_e2e:
        hp = host_frames.numpy()
        for _ in range(2):
            ctx.estimate_batch(hp)
        barrier()
        e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        steps_e2e = max(2, min(steps, 3))
        tw0 = time.time()
        e2[0].record(stream)
        t0 = time.perf_counter()
        for _ in range(steps_e2e):
            for _b in range(bps):
                r = ctx.estimate_batch(hp)
        e2[1].record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tw1 = time.time()
        ms_mine = max(e2[0].elapsed_time(e2[1]), wall * 1e3) / steps_e2e
        t = torch.tensor([ms_mine], dtype=torch.float64, device=dev)
        per_rank = None
        if world > 1:
            allms = torch.empty(world, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allms, t)
            per_rank = [bps * B * W * H / (float(m) * 1e-3) / 1e9 for m in allms.tolist()]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
        e2e = {"value": world * B * bps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": bps * B * W * H,
               "d2h_bytes_per_step": bps * B * rec, "ms_per_step": ms_e2e, "steps": steps_e2e,
               "h2d_gbs": bps * B * W * H / (ms_e2e * 1e-3) / 1e9,      # per GPU; PCIe-bound: compare with the box's pinned H2D copy rate (55.6 GB/s measured, tests/probes/pcie_probe.py)
               "h2d_gbs_per_rank": per_rank, "h2d_gbs_aggregate": sum(per_rank) if per_rank else None, "host_placement": bind_info,
               "clocks": sampler.window(tw0, tw1),
               "api": "mpe_estimate_batch (pinned host frames -> host mpe_result records)"}
        assert int(results_to_arrays(r)["updated"].sum()) == n_updated
    out = dict(scene=scene, kt=kt, value=value, ms_step=ms_step, clocks=clocks, e2e=e2e, launches=launches_timed, host_enqueue_ms_per_batch=host_enqueue_ms_per_batch, ms_step_per_rank=ms_step_per_rank, n_ctx=n_ctx, n_updated=n_updated,
               gather=gather_info, oracle_check=oracle_check, fp64_peak=fp64_peak, window=(t_load0, t_load1))
    if do_cpu and rank == 0:
        fps, n, dt = cpu_single_thread(scene, max_seconds=cpu_seconds)
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": f"{n} frames of the same batch in {dt:.1f} s, one thread (cv2 4.13 findLeds + C++ oracle of the pose path)",
                               "host_cpus": os.cpu_count(), "host_cpus_usable": usable_cores()}
    for c in ctxs:
        c.close()
    del dev_frames, host_frames
    torch.cuda.empty_cache()
    return out


def kernel_table(m, B, W, H, leds, peak, ops):
    names = ["scan (K1a)", "extract_blobs (K1b)", "p3p_sweep (K2)", "check+refine (K3)", "blur_tiles (K1c)"]
    kt = m["kt"]
    ksum = sum(kt)
    k1_gbs = B * W * H / (kt[0] * 1e-3) / 1e9 if kt[0] > 0 else 0.0
    kernels = [{"name": n, "ms": ms, "share_of_kernel_time": ms / ksum if ksum > 0 else None} for n, ms in zip(names, kt)]
    kernels[0].update({"bound": "hbm", "achieved_gbs": k1_gbs, "frac": k1_gbs / peak})
    kernels[1].update({"bound": "latency (sparse contour tracing)"})
    n_prob = leds * (leds - 1) * (leds - 2) // 6 * leds * (leds - 1) * (leds - 2)
    kernels[2].update({"bound": "fp64 alu", "p3p_problems_per_s": B * n_prob / (kt[2] * 1e-3) if kt[2] > 0 else None})
    kernels[3].update({"bound": "fp64 alu / latency (dependent GN chain)"})
    kernels[4].update({"bound": "latency (sparse exact blur of hot tiles)"})
    fp64 = None
    if ops and m.get("fp64_peak"):
        f_init = ops["initialise_flops_per_frame"]
        t_k2 = B * f_init / (kt[2] * 1e-3) / 1e12 if kt[2] > 0 else 0.0
        kernels[2].update({"flops_per_frame": f_init, "achieved_tflops": t_k2, "frac": t_k2 / m["fp64_peak"]})
        if ops.get("optimise_flops_per_frame"):
            t_k3 = B * ops["optimise_flops_per_frame"] / (kt[3] * 1e-3) / 1e12 if kt[3] > 0 else 0.0
            kernels[3].update({"flops_per_frame_optimise_only": ops["optimise_flops_per_frame"], "achieved_tflops": t_k3, "frac": t_k3 / m["fp64_peak"]})
        fp64 = {"kernel": "p3p_sweep (K2: tier-1 pre-test + exact P3P solve + scoring; initialise() of the reference)", "bound": "fp64", "achieved": t_k2,
                "peak": m["fp64_peak"], "unit": "TFLOP/s", "frac": t_k2 / m["fp64_peak"],
                "peak_source": "measured here: mpe_probe_fp64_peak (independent DFMA chains, FMA = 2 flops); without FMA contraction, as the reference's arithmetic "
                               "is compiled (-fmad=false), the ceiling for separate multiplies and adds is half of it",
                "algorithmic_flops_per_launch": B * f_init, "launch_ms": kt[2], "op_count": ops,
                "note": "ALGORITHMIC flops = what the reference's initialise() performs on these frames (exact count); the kernel does fewer: tier 1 rules "
                        "out ~95 % of the P3P problems with a quarter of the operations, so the fraction measures time-to-solution against the FP64 peak"}
    return kernels, names[int(np.argmax(kt))], k1_gbs, ksum, fp64


def run_cold(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    bind_info = bind_near_gpu(local_rank) if not args.no_bind else {"bound": False}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one poller per node (local rank 0 watches every GPU of the job); the other ranks do not touch nvidia-smi
    sampler = ClockSampler(0, n_gpus=max(world, 1), enabled=(local_rank == 0)) if world > 1 else ClockSampler(local_rank)
    sampler.start()                      # started early: nvidia-smi needs ~1 s before its first sample
    W, H, B, bps = args.width, args.height, args.batch, args.batches_per_step
    m = cold_measure(args, local_rank, world, rank, dist if world > 1 else None, sampler, B, bps, args.steps, args.warmup, args.leds, W, H, args.seed,
                     args.contexts, not args.no_e2e, not args.no_cpu, args.cpu_seconds, 256, bind_info=bind_info)
    if rank == 0:
        peak, peak_src = load_peaks()
        ops = k2_op_counts(m["scene"]) if not args.no_cpu else None
        kernels, dominant, k1_gbs, ksum, fp64 = kernel_table(m, B, W, H, args.leds, peak, ops)
        line = {
            "metric": METRIC if (W, H, args.leds) == (WIDTH, HEIGHT, N_LEDS) else f"frames/sec ({W}x{H}, {args.leds} LEDs, cold)",
            "value": m["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": m["ms_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (findLeds, fixed point) + f64 (P3P, Gauss-Newton)",
            "data": "synthetic",
            "config": make_config(args),
            "parallelism": f"frames sharded over {world} GPU(s), NCCL all-gather of the result records per batch" if world > 1 else "1 GPU",
            "batches_in_flight": m["n_ctx"], "contour_frames_per_warp": int(os.environ.get("MPE_K1B_POOL", "4")) or 1, "frames_with_pose_per_batch": m["n_updated"], "oracle_check": m["oracle_check"],
            "parity_note": "the CPU oracle these results are checked against is pinned to the UNMODIFIED reference sources built in oracle/_ref (stand-in Eigen): "
                           "discrete outputs and Gauss-Newton iteration counts equal; what no build here can pin is real Eigen's last-ulp rounding, and the "
                           "exit test of optimisePose (1e-13) sits at that floor, so +-1 iteration against a real-Eigen binary cannot be excluded",
            "clocks": m["clocks"],
            "e2e": m["e2e"],
            "gpu_launches": m["launches"],
            "host_enqueue_ms_per_batch": m["host_enqueue_ms_per_batch"],
            "ms_per_step_per_rank": m["ms_step_per_rank"],
            "gather": m["gather"],
            "roofline": {"kernel": "scan_kernel (K1a: TMA-streamed threshold scan of every ROI byte; findLeds hot loop)", "bound": "hbm", "achieved": k1_gbs, "peak": peak, "unit": "GB/s",
                         "frac": k1_gbs / peak, "peak_source": peak_src, "traffic": load_traffic(B, W, H),
                         "algorithmic_bytes_per_launch": B * W * H, "launch_ms": m["kt"][0]},
            "roofline_fp64": fp64,
            "dominant_kernel": dominant,
            "kernel_times_note": "stage times from a separate pass on one context with CUDA events around every stage and the record gather switched off "
                                 "(sum = %.3f ms per batch); in the timed steps %d batches are in flight on %d streams, so ms_per_step / batches_per_step can be "
                                 "smaller than that sum" % (ksum, m["n_ctx"], m["n_ctx"]),
            "kernels": kernels,
        }
        if "cpu_baseline" in m:
            line["cpu_baseline"] = m["cpu_baseline"]
        # ---- the other BASELINE configurations, same run, own clock windows (1 GPU only: the scaling runs stay short)
        if world == 1 and not args.no_extras:
            extra = {}
            a2 = argparse.Namespace(**vars(args))
            try:
                extra["tracking"] = tracking_measure(a2, local_rank, B, 256, 200, 3, sampler, cpu=not args.no_cpu, e2e=not args.no_e2e)
            except Exception as e:
                extra["tracking"] = {"error": repr(e)[:200]}
            try:
                extra["single_camera_latency"] = latency_measure(a2, sampler, n_frames=200, sweep=(1, 8, 64), variants=False)
            except Exception as e:
                extra["single_camera_latency"] = {"error": repr(e)[:200]}
            for key, (w2, h2, l2, b2, bps2, secs) in {"config3_1920x1080": (1920, 1080, 5, 768, 8, 6.0), "config4_8_leds": (752, 480, 8, 1024, 2, 6.0)}.items():
                try:
                    m2 = cold_measure(a2, local_rank, 1, 0, None, sampler, b2, bps2, 20, 3, l2, w2, h2, args.seed, args.contexts, not args.no_e2e, not args.no_cpu, secs, 8,
                                      do_gather=False, bind_info=bind_info)
                    ops2 = k2_op_counts(m2["scene"], n_frames=(4 if l2 < 8 else 2)) if not args.no_cpu else None
                    k2t, dom2, g2, _, fp2 = kernel_table(m2, b2, w2, h2, l2, peak, ops2)
                    extra[key] = {"metric": f"frames/sec ({w2}x{h2}, {l2} LEDs, cold)", "value": m2["value"], "unit": "frames/s", "ms_per_step": m2["ms_step"],
                                  "config": {"width": w2, "height": h2, "leds": l2, "mode": "cold", "batch": b2, "batches_per_step": bps2}, "steps": 20, "warmup": 3,
                                  "clocks": m2["clocks"], "e2e": m2["e2e"], "roofline": {"kernel": "scan_kernel (K1a)", "bound": "hbm", "achieved": g2, "peak": peak, "unit": "GB/s", "frac": g2 / peak,
                                                                                          "launch_ms": m2["kt"][0], "algorithmic_bytes_per_launch": b2 * w2 * h2},
                                  "roofline_fp64": fp2, "dominant_kernel": dom2, "kernels": k2t, "cpu_baseline": m2.get("cpu_baseline"), "oracle_check": m2["oracle_check"],
                                  "frames_with_pose_per_batch": m2["n_updated"]}
                except Exception as e:
                    extra[key] = {"error": repr(e)[:200]}
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # Several ranks: the record gather (an NCCL kernel) shares each GPU with the pipeline.  The four-frames-per-warp contour kernel
    # (long-running CTAs on every SM) then keeps NCCL's CTAs waiting and the collective resident for ~0.5 ms per batch instead of
    # 0.05 ms: measured on 2 x B200 10.63 M frames/s with it, 11.01 M without.  Alone on the GPU it is the faster choice (5.85 M
    # against 5.59 M frames/s).  The library reads the knob once, at its first contour launch.
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MPE_K1B_POOL", "0")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="frames per GPU per batch (one launch sequence)")
    ap.add_argument("--batches-per-step", type=int, default=16, help="batches per step: stretches the timed region (20 steps x 16 batches ~ 0.5 s)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--leds", type=int, default=N_LEDS)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--mode", default="cold", choices=["cold", "tracking", "latency"],
                    help="cold (headline): every frame runs the full pipeline; tracking: device-resident streams with ROI search (with --gpus N: config 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs and the op counts")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configurations (tracking, latency, 1080p, 8 LEDs)")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the CPUs next to its GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-gather", action="store_true", help="N > 1: leave the NCCL gather of the records out (diagnostic)")
    ap.add_argument("--gather-slots", type=int, default=2, help="N > 1: depth of the ring of gather buffers")
    ap.add_argument("--contexts", type=int, default=2, help="batches in flight in the device-resident measurement (one context + stream each)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.mode == "tracking":
        run_tracking(args)
    elif args.mode == "latency":
        run_latency(args)
    else:
        run_cold(args)


if __name__ == "__main__":
    main()
