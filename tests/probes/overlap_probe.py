"""Probe: does splitting a batch over several contexts/streams (kernels of different chunks overlapping) raise throughput?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth
B = 8192
sc = synth.make_cold_scene(512, n_leds=5, seed=7)
frames = torch.from_numpy(sc.frames).cuda().repeat(B // 512, 1, 1).contiguous()
W, H = 752, 480
for nctx in (1, 2, 4, 8):
    n = B // nctx
    ctxs = []
    for i in range(nctx):
        c = mpe.Context(0, n, W, H); c.set_camera(sc.K, sc.D); c.set_params(sc.params); c.set_markers(sc.markers); ctxs.append(c)
    def step():
        for i, c in enumerate(ctxs):
            c.estimate_batch_device_async(frames[i * n].data_ptr(), W, W * H, W, H, n)
    for _ in range(3): step()
    torch.cuda.synchronize()
    for c in ctxs: c.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step()
    for c in ctxs: c.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{nctx} contexts x {n} frames: {dt*1e3:.3f} ms/step -> {B/dt/1e6:.3f} M frames/s")
    for c in ctxs: c.close()
