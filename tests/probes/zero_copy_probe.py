"""Probe: does the tracking step work when the `frames_device` pointer is PINNED HOST memory (UVA), i.e. can the K1 TMA tile
loads and the blur kernel's pixel reads go straight over PCIe so that only ROI bytes leave the host?  Compares the records with
the device-buffer run and times both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth
from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays

S, T = 256, 12
seqs = [synth.make_stream_scene(T, n_leds=5, seed=500 + s) for s in range(16)]
W, H = seqs[0].width, seqs[0].height
frames = np.stack([np.stack([seqs[s % 16].frames[t] for s in range(S)]) for t in range(T)])   # T x S x H x W
host = torch.from_numpy(frames).pin_memory()
dev = host.cuda()
times = [np.array([seqs[s % 16].times[t] for s in range(S)]) for t in range(T)]
out = {}
for label, buf in (("device", dev), ("pinned_host", host)):
    ctx = mpe.Context(0, S, W, H)
    ctx.set_camera(seqs[0].K, seqs[0].D); ctx.set_params(seqs[0].params); ctx.set_markers(seqs[0].markers)
    ctx.streams_reset(S)
    recs, dts = [], []
    for t in range(T):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ctx.streams_step_device(buf[t].data_ptr(), W, W * H, W, H, times[t])
        dts.append(time.perf_counter() - t0)
        recs.append(results_to_arrays(r).copy())
    out[label] = recs
    print(label, "ms/step", [round(x * 1e3, 3) for x in dts], "updated last", int(recs[-1]["updated"].sum()), flush=True)
    ctx.close()
same = all(a.tobytes() == b.tobytes() for a, b in zip(out["device"], out["pinned_host"]))
print("records identical:", same)
