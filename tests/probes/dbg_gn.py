import sys, numpy as np
sys.path.insert(0, '.')
import rpg_monocular_pose_estimator_b200 as mpe
from rpg_monocular_pose_estimator_b200 import synth
from oracle import pose_oracle
from tests.helpers import oracle_find_leds
sc = synth.make_cold_scene(1, n_leds=5, seed=321)
ctx = mpe.Context(0, 4, 752, 480)
ctx.set_camera(sc.K, sc.D); ctx.set_params(sc.params); ctx.set_markers(sc.markers)
det, _ = oracle_find_leds(sc.frames[0], (0,0,752,480), sc.params, sc.K, sc.D)
est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
est.set_image_points(det); assert est.initialise() == 1
corr = est.correspondences(); T0 = est.predicted_pose()
# expected A, b at T0
K = sc.K; fx, fy = K[0,0], K[1,1]
A = np.zeros((6,6)); b = np.zeros(6)
for led, d in corr:
    X = np.append(sc.markers[led-1], 1.0); pc = T0 @ X
    uv = (K @ pc[:3]); uv = uv[:2]/uv[2]
    e = det[d-1] - uv
    x,y,z = pc[:3]; z2 = z*z
    J = np.array([[1/z*fx, 0, -x/z2*fx, -x*y/z2*fx, (1+x*x/z2)*fx, -y/z*fx],[0, 1/z*fy, -y/z2*fy, -(1+y*y/z2)*fy, x*y/z2*fy, x/z*fy]])
    A += J.T@J; b += J.T@e
np.set_printoptions(precision=9, linewidth=200)
print("expected A\n", A, "\nexpected b", b, "\nexpected dT", np.linalg.solve(A,b))
pose2, cov, it = ctx.optimise_pose(det, corr, T0)
print("iters", it)
