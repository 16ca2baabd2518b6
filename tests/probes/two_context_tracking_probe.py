"""Probe: tracking mode with the streams split over 1, 2 or 4 contexts whose CUDA-graph replays run on separate CUDA streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth

S, Sd, T, W, H = 8192, 256, 40, 752, 480
seqs = [synth.make_stream_scene(T, n_leds=5, seed=12345 + 17 * s) for s in range(Sd)]
buf = torch.from_numpy(np.stack([f for sc in seqs for f in sc.frames])).cuda()
for n_ctx in (1, 2, 4, 1, 2):
    Sk = S // n_ctx
    parts = []
    for k in range(n_ctx):
        c = mpe.Context(0, Sk, W, H); c.set_camera(seqs[0].K, seqs[0].D); c.set_params(seqs[0].params); c.set_markers(seqs[0].markers)
        st = torch.cuda.Stream(); c.set_stream(st.cuda_stream); c.streams_reset(Sk)
        base = ((torch.arange(Sk, dtype=torch.int32, device="cuda") + k * Sk) % Sd) * T
        fmap = base.clone(); c.streams_set_frame_map(fmap.data_ptr(), Sd * T)
        parts.append((c, st, base, fmap))
    def step(t):
        for c, st, base, fmap in parts:
            with torch.cuda.stream(st):
                fmap.copy_(base + t)
            c.streams_step_device(buf.data_ptr(), W, W * H, W, H, np.full(Sk, t / 60.0), fetch=False)
    for t in range(12): step(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(12, 36): step(t)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 24 * 1e3
    upd = sum(int(np.frombuffer(c.fetch_results(Sk), dtype=np.uint8).reshape(Sk, -1)[:, 0].sum()) for c, *_ in parts)
    print(n_ctx, "context(s):", round(ms, 3), "ms per step of", S, "streams ->", round(S / ms / 1e3, 2), "M frames/s; updated", upd, flush=True)
    for c, *_ in parts: c.close()
