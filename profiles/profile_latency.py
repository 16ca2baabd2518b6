"""ncu workload: a single stream advanced frame by frame with plain launches (graph replay off), so that every kernel of
one tracking step shows up as its own launch.  Used with: ncu --metrics gpu__time_duration.sum --clock-control none ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = synth.make_stream_scene(T, n_leds=5, seed=12345)
ctx = mpe.Context(0, 1, sc.width, sc.height)
ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
ctx.set_graph_replay(False)
ctx.streams_reset(1)
for t in range(T):
    r = ctx.streams_step(sc.frames[t][None], [sc.times[t]])
    print(t, r[0].updated, r[0].gn_iters, r[0].n_det)
ctx.close()
