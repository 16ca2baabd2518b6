"""The C++ class shim (include/monocular_pose_estimator_b200/shim.h) driven like MPENode drives the reference library must
reproduce the oracle's estimateBodyPose sequence (cold start + ROI tracking)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from oracle import pose_oracle
from tests.helpers import pose_error

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["device_loop", "stage"])
@pytest.mark.parametrize("types", ["standin", "real_types"])
def test_cpp_shim_tracking_sequence(tmp_path, mode, types):
    """types = real_types: the shim compiled with the reference's own type names (Eigen::Matrix..., cv::Mat; against the Eigen /
    OpenCV stand-ins of oracle/) and driven by a copy of MPENode's call sequence incl. augmentImage."""
    exe = os.path.join(ROOT, "build", "shim_demo" if types == "standin" else "shim_real_types_demo")
    if not os.path.exists(exe):
        subprocess.check_call(["python", "-c", "import __graft_entry__ as g; g.build_shim_demos()"], cwd=ROOT)
    sc = synth.make_stream_scene(20, n_leds=5, seed=21)
    p = sc.params
    scene = tmp_path / "scene.bin"
    with open(scene, "wb") as f:
        f.write(struct.pack("4i", len(sc.frames), sc.width, sc.height, len(sc.markers)))
        f.write(np.ascontiguousarray(sc.K, np.float64).tobytes()); f.write(np.ascontiguousarray(sc.D[:5], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc.markers, np.float64).tobytes())
        f.write(np.array([p.threshold_value, p.gaussian_sigma, p.min_blob_area, p.max_blob_area, p.max_width_height_distortion,
                          p.max_circular_distortion, p.back_projection_pixel_tolerance, p.nearest_neighbour_pixel_tolerance,
                          p.certainty_threshold, p.valid_correspondence_threshold, p.roi_border_thickness], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc.times, np.float64).tobytes()); f.write(sc.frames.tobytes())
    out = subprocess.run([exe, str(scene)] + (["stage"] if mode == "stage" else []), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(lines) == len(sc.frames)
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    n_ok = 0
    for fi, l in enumerate(lines):
        upd = est.estimate_body_pose(sc.frames[fi], sc.times[fi])
        assert int(l[1]) == int(upd), fi
        assert tuple(int(v) for v in l[2:6]) == tuple(est.region_of_interest), fi
        if upd:
            n_ok += 1
            T = np.array([float(v) for v in l[7:23]]).reshape(4, 4)
            dt, dr = pose_error(T, est.predicted_pose())
            assert dt < 1e-6 and dr < 1e-6, (fi, dt, dr)
            assert int(l[6]) == est.gn_iterations()
    assert n_ok >= 18


def test_unmodified_ros_node_on_the_shim(tmp_path):
    """build/mpenode_on_shim = the reference's ROS node source, unmodified, on the class shim (tests/cpp/mpenode_harness.cpp plays
    roscore: parameter server, CameraInfo, dynamic_reconfigure, one sensor_msgs/Image per frame).  What the node publishes must be the
    oracle's estimateBodyPose sequence; position / orientation / covariance[0] of the PoseWithCovarianceStamped must be the pose's.
    (First run on a B200: profiles/mpenode_on_shim_r02_output.txt — 20 of 20 frames published, poses equal to 1e-15.)"""
    from scipy.spatial.transform import Rotation
    exe = os.path.join(ROOT, "build", "mpenode_on_shim")
    if not os.path.exists(exe):
        pytest.skip("build/mpenode_on_shim is built where /root/reference exists (__graft_entry__.build_mpenode_on_shim)")
    sc = synth.make_stream_scene(20, n_leds=5, seed=21)
    p = sc.params
    scene = tmp_path / "scene.bin"
    with open(scene, "wb") as f:
        f.write(struct.pack("4i", len(sc.frames), sc.width, sc.height, len(sc.markers)))
        f.write(np.ascontiguousarray(sc.K, np.float64).tobytes()); f.write(np.ascontiguousarray(sc.D[:5], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc.markers, np.float64).tobytes())
        f.write(np.array([p.threshold_value, p.gaussian_sigma, p.min_blob_area, p.max_blob_area, p.max_width_height_distortion,
                          p.max_circular_distortion, p.back_projection_pixel_tolerance, p.nearest_neighbour_pixel_tolerance,
                          p.certainty_threshold, p.valid_correspondence_threshold, p.roi_border_thickness], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc.times, np.float64).tobytes()); f.write(sc.frames.tobytes())
    out = subprocess.run([exe, str(scene)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(lines) == len(sc.frames)
    est = pose_oracle.PoseEstimatorOracle(sc.K, sc.D, sc.markers, sc.params)
    n_ok = 0
    for fi, l in enumerate(lines):
        upd = est.estimate_body_pose(sc.frames[fi], sc.times[fi])
        assert int(l[1]) == int(upd), fi                               # published <=> pose updated
        assert tuple(int(v) for v in l[2:6]) == tuple(est.region_of_interest), fi
        if upd:
            n_ok += 1
            T = np.array([float(v) for v in l[7:23]]).reshape(4, 4)
            dt, dr = pose_error(T, est.predicted_pose())
            assert dt < 1e-6 and dr < 1e-6, (fi, dt, dr)
            assert int(l[6]) == est.gn_iterations()
            msg = [float(v) for v in l[24:32]]                         # after the "|": position xyz, orientation xyzw, covariance[0]
            assert np.allclose(msg[:3], T[:3, 3], rtol=0, atol=0)
            q = Rotation.from_matrix(T[:3, :3]).as_quat()
            assert min(np.abs(q - msg[3:7]).max(), np.abs(q + msg[3:7]).max()) < 1e-12
    assert n_ok >= 18
