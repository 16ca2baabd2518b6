"""Diagnostic: where one camera's tracking step spends its time.  MPE_STEP_TRACE=1 prints CUDA-event stage times of plain launches;
the graph-replay wall clock is measured next to it."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

W, H = 752, 480
T = 140
sc = synth.make_stream_scene(T, n_leds=5, width=W, height=H, seed=12345)
slot = torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy()
for graphs in (False, True):
    ctx = mpe.Context(0, 1, W, H)
    ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
    ctx.set_graph_replay(graphs)
    ctx.streams_reset(1)
    res = (mpe.MpeResult * 1)()
    tarr = np.zeros(1)
    tp = tarr.ctypes.data_as(C.POINTER(C.c_double))
    ptr = C.c_void_p(slot.ctypes.data)
    lat = []
    for t in range(T):
        tarr[0] = sc.times[t]
        slot[:] = sc.frames[t]
        t0 = time.perf_counter()
        rc = ctx.L.mpe_streams_step(ctx.h, ptr, W, W * H, W, H, 1, tp, res)
        t1 = time.perf_counter()
        assert rc == 0
        if t >= 40:
            lat.append((t1 - t0) * 1e6)
    print("graphs", graphs, "p50 %.1f us  p10 %.1f  p90 %.1f" % tuple(np.percentile(lat, [50, 10, 90])), flush=True)
    ctx.close()
