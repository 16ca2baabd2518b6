"""K2 sweep: results and kernel time for the three filter modes (0 exact, 1 filter behind the solve, 2 tier 1)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth
from rpg_monocular_pose_estimator_b200.pose_estimator import results_to_arrays

def run(n_leds, B, modes=(0, 1, 2)):
    nd = min(B, 256)
    sc = synth.make_cold_scene(nd, n_leds=n_leds, seed=5)
    ctx = mpe.Context(0, B, 752, 480)
    ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
    frames = np.ascontiguousarray(np.tile(sc.frames, ((B + nd - 1) // nd, 1, 1))[:B])
    d = torch.from_numpy(frames).cuda()
    ctx.enable_kernel_timing(True)
    out = {}
    for mode in modes:
        ctx.set_k2_filter(mode)
        for _ in range(3):
            res = results_to_arrays(ctx.estimate_batch_device(d.data_ptr(), 752, 752 * 480, 752, 480, B))
        t = ctx.kernel_times_ms()
        out[mode] = res
        print(n_leds, B, "mode", mode, "kernel ms [scan, extract, sweep, refine, blur]", [round(float(x), 4) for x in t],
              "updated", int(res["updated"].sum()), flush=True)
    for mode in modes[1:]:
        a, b = out[mode], out[modes[0]]
        same = all(np.array_equal(a[k], b[k]) for k in ("updated", "n_corr", "gn_iters", "corr", "init_ok")) and \
            np.array_equal(a["pose"][a["updated"] == 1], b["pose"][b["updated"] == 1])
        print("mode", mode, "records identical to mode", modes[0], ":", same, flush=True)
    ctx.close()

if __name__ == "__main__":
    run(5, 8192)
    run(8, 1024)
    run(4, 8192)
