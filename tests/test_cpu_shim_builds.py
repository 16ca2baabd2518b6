"""The C++ class shim must compile and link against the C-ABI library in both of its modes: with its own stand-in types and with
the reference's type names (MPE_SHIM_REAL_TYPES: Eigen::Matrix<...>, cv::Mat, cv::Rect — here against the Eigen / OpenCV stand-ins
under oracle/, since the real headers are not installed).  The real-types driver is a copy of MPENode's call sequence
(monocular_pose_estimator/src/monocular_pose_estimator.cpp:103-236) including augmentImage and all 15 datatypes.h typedefs.
Running them needs a GPU (tests/test_gpu_cpp_shim.py)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_demos_compile_and_link():
    import __graft_entry__ as g
    for exe in ("shim_demo", "shim_real_types_demo"):
        p = os.path.join(ROOT, "build", exe)
        if os.path.exists(p):
            os.unlink(p)
    g.build_shim_demos()
    for exe in ("shim_demo", "shim_real_types_demo"):
        assert os.path.exists(os.path.join(ROOT, "build", exe))
    # without a device the shim must fail loudly, not fall back to anything
    out = subprocess.run([os.path.join(ROOT, "build", "shim_real_types_demo")], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr
