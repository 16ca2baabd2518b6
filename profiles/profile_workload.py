"""Workload driven under ncu (never a bench value): W warm-up + K steps of the cold path on device-resident frames."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--leds", type=int, default=5)
ap.add_argument("--width", type=int, default=752)
ap.add_argument("--height", type=int, default=480)
a = ap.parse_args()
n_distinct = min(a.batch, 512)
sc = synth.make_cold_scene(n_distinct, n_leds=a.leds, width=a.width, height=a.height, seed=7)
frames = torch.from_numpy(sc.frames).cuda()
frames = frames.repeat((a.batch + n_distinct - 1) // n_distinct, 1, 1)[:a.batch].contiguous()
ctx = mpe.Context(0, a.batch, a.width, a.height)
ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
for _ in range(a.warmup + a.steps):
    ctx.estimate_batch_device_async(frames.data_ptr(), a.width, a.width * a.height, a.width, a.height, a.batch)
ctx.synchronize()
print("done", ctx.launch_count())
