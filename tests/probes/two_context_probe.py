"""Probe: do two contexts on two streams, alternating whole cold batches, beat one context (the tail kernels of batch i —
thread-per-frame check/refine, < 2 warps per SM — could overlap the HBM-bound scan of batch i+1)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

B, W, H = 8192, 752, 480
sc = synth.make_cold_scene(512, n_leds=5, seed=3)
frames = torch.from_numpy(sc.frames).cuda().repeat(B // 512, 1, 1).contiguous()
def make():
    c = mpe.Context(0, B, W, H); c.set_camera(sc.K, sc.D); c.set_params(sc.params); c.set_markers(sc.markers)
    s = torch.cuda.Stream(); c.set_stream(s.cuda_stream); return c, s
ctxs = [make(), make()]
def run(n_ctx, steps=20):
    for c, s in ctxs[:n_ctx]:
        c.estimate_batch_device_async(frames.data_ptr(), W, W * H, W, H, B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        c, s = ctxs[i % n_ctx]
        c.estimate_batch_device_async(frames.data_ptr(), W, W * H, W, H, B)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for n in (1, 2, 1, 2):
    print(n, "context(s):", round(run(n), 3), "ms per batch of", B, flush=True)
