"""ncu workload: S device-resident tracking streams, plain launches (graph replay off), a few settled steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Sd, T = 128, 12
seqs = [synth.make_stream_scene(T, n_leds=5, seed=12345 + 17 * s) for s in range(Sd)]
W, H = seqs[0].width, seqs[0].height
buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).cuda()
ctx = mpe.Context(0, S, W, H)
ctx.set_camera(seqs[0].K, seqs[0].D); ctx.set_params(seqs[0].params); ctx.set_markers(seqs[0].markers)
ctx.set_graph_replay(False)
ctx.streams_reset(S)
base = (torch.arange(S, dtype=torch.int32, device="cuda") % Sd) * T
fmap = base.clone()
ctx.streams_set_frame_map(fmap.data_ptr(), Sd * T)
for t in range(T):
    fmap.copy_(base + t)
    ctx.streams_step_device(buf.data_ptr(), W, W * H, W, H, np.full(S, t / 60.0), fetch=False)
ctx.synchronize()
print("done", ctx.launch_count())
