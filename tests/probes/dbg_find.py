import sys, numpy as np
sys.path.insert(0, '.')
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth
from tests.helpers import oracle_find_leds
sc = synth.make_cold_scene(1, n_leds=5, seed=1)
ctx = mpe.Context(0, 4, 752, 480)
ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params)
px, ce, fl = ctx.find_leds(sc.frames[0], (0, 0, 752, 480))
print("gpu", ce, fl)
print("ora", oracle_find_leds(sc.frames[0], (0, 0, 752, 480), sc.params, sc.K, sc.D)[1])
