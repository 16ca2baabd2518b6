"""ctypes binding of include/mpe_b200.h.  No fallback: a missing library is an ImportError-like failure."""
from __future__ import annotations

import ctypes as C
import os

MPE_MAX_LEDS = 16
MPE_MAX_DET = 16
MPE_MAX_BLOBS = 64
MPE_MAX_DIST = 12

MPE_F_BLOB_OVERFLOW = 1
MPE_F_TRACE_ABORT = 2
MPE_F_TOO_MANY_DET = 4
MPE_F_INITIALISED = 8
MPE_F_FULL_IMAGE_RETRY = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpe_b200.so")
if os.environ.get("MPE_B200_LIB"):      # A/B experiments: another in-tree build of the same library
    LIB_PATH = os.environ["MPE_B200_LIB"]


class MpeError(RuntimeError):
    pass


class MpeParams(C.Structure):
    _fields_ = [("threshold_value", C.c_int32), ("roi_border_thickness", C.c_int32), ("gaussian_sigma", C.c_double),
                ("min_blob_area", C.c_double), ("max_blob_area", C.c_double), ("max_width_height_distortion", C.c_double),
                ("max_circular_distortion", C.c_double), ("back_projection_pixel_tolerance", C.c_double),
                ("nearest_neighbour_pixel_tolerance", C.c_double), ("certainty_threshold", C.c_double),
                ("valid_correspondence_threshold", C.c_double)]


class MpeRect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class MpeResult(C.Structure):
    _fields_ = [("updated", C.c_int32), ("n_det", C.c_int32), ("n_corr", C.c_int32), ("gn_iters", C.c_int32),
                ("flags", C.c_int32), ("init_ok", C.c_int32), ("roi", MpeRect), ("pose", C.c_double * 16),
                ("cov", C.c_double * 36), ("corr", C.c_uint32 * (2 * MPE_MAX_LEDS)), ("det", C.c_double * (2 * MPE_MAX_DET)),
                ("centers", C.c_float * (2 * MPE_MAX_DET))]


EXPORTS = [
    "mpe_create", "mpe_destroy", "mpe_last_error", "mpe_set_stream", "mpe_set_camera", "mpe_set_markers", "mpe_set_params",
    "mpe_set_histogram_threshold", "mpe_get_histogram_threshold", "mpe_find_leds", "mpe_initialise",
    "mpe_check_correspondences", "mpe_optimise_pose", "mpe_p3p_compute_poses", "mpe_estimate_batch",
    "mpe_estimate_batch_device", "mpe_estimate_batch_device_async", "mpe_fetch_results", "mpe_synchronize", "mpe_copy_poses_device",
    "mpe_streams_reset", "mpe_streams_set_frame_map", "mpe_streams_step_device", "mpe_streams_step", "mpe_set_graph_replay", "mpe_set_ingest_mode", "mpe_get_ingest_stats", "mpe_set_k2_filter", "mpe_copy_results_device", "mpe_probe_fp64_peak", "mpe_debug_gaussian_taps", "mpe_host_predict_pose", "mpe_host_project_markers", "mpe_host_determine_roi", "mpe_host_exponential_map", "mpe_host_logarithm_map",
    "mpe_enable_kernel_timing", "mpe_get_kernel_times",
    "mpe_kernel_launch_count", "mpe_pose_to_message",
]

_LIB = None


def load_library():
    """Loads libmpe_b200.so (built in-tree by __graft_entry__.build() / csrc/build.sh)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise MpeError(f"native library missing: {LIB_PATH} — run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, fp, ip, up = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint32)
    sig = {
        "mpe_create": ([C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "mpe_destroy": ([vp], None),
        "mpe_last_error": ([vp], C.c_char_p),
        "mpe_set_stream": ([vp, vp], C.c_int),
        "mpe_set_camera": ([vp, dp, dp, C.c_int], C.c_int),
        "mpe_set_markers": ([vp, dp, C.c_int], C.c_int),
        "mpe_set_params": ([vp, C.POINTER(MpeParams)], C.c_int),
        "mpe_set_histogram_threshold": ([vp, C.c_uint32], C.c_int),
        "mpe_get_histogram_threshold": ([vp], C.c_uint32),
        "mpe_find_leds": ([vp, vp, C.c_int, C.c_int, C.c_int, MpeRect, dp, fp, ip, ip], C.c_int),
        "mpe_initialise": ([vp, dp, C.c_int, up, up, ip, dp, ip], C.c_int),
        "mpe_check_correspondences": ([vp, dp, C.c_int, up, C.c_int, dp, ip], C.c_int),
        "mpe_optimise_pose": ([vp, dp, C.c_int, up, C.c_int, dp, dp, ip], C.c_int),
        "mpe_p3p_compute_poses": ([vp, dp, dp, C.c_int, dp, ip], C.c_int),
        "mpe_estimate_batch": ([vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(MpeResult)], C.c_int),
        "mpe_estimate_batch_device": ([vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(MpeResult)], C.c_int),
        "mpe_estimate_batch_device_async": ([vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int], C.c_int),
        "mpe_fetch_results": ([vp, C.c_int, C.POINTER(MpeResult)], C.c_int),
        "mpe_synchronize": ([vp], C.c_int),
        "mpe_copy_poses_device": ([vp, C.c_int, vp], C.c_int),
        "mpe_streams_reset": ([vp, C.c_int], C.c_int),
        "mpe_streams_set_frame_map": ([vp, vp, C.c_int], C.c_int),
        "mpe_streams_step_device": ([vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, dp, C.POINTER(MpeResult)], C.c_int),
        "mpe_streams_step": ([vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, dp, C.POINTER(MpeResult)], C.c_int),
        "mpe_set_graph_replay": ([vp, C.c_int], C.c_int),
        "mpe_set_ingest_mode": ([vp, C.c_int], C.c_int),
        "mpe_set_k2_filter": ([vp, C.c_int], C.c_int),
        "mpe_copy_results_device": ([vp, C.c_int, vp], C.c_int),
        "mpe_probe_fp64_peak": ([vp, C.POINTER(C.c_double)], C.c_int),
        "mpe_host_predict_pose": ([dp, dp, C.c_double, C.c_double, C.c_double, dp], C.c_int),
        "mpe_host_project_markers": ([dp, dp, dp, C.c_int, dp], C.c_int),
        "mpe_host_determine_roi": ([dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.POINTER(MpeRect)], C.c_int),
        "mpe_host_exponential_map": ([dp, dp], C.c_int),
        "mpe_host_logarithm_map": ([dp, dp], C.c_int),
        "mpe_debug_gaussian_taps": ([C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_int], C.c_int),
        "mpe_get_ingest_stats": ([vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)], C.c_int),
        "mpe_enable_kernel_timing": ([vp, C.c_int], C.c_int),
        "mpe_get_kernel_times": ([vp, fp], C.c_int),
        "mpe_kernel_launch_count": ([vp], C.c_longlong),
        "mpe_pose_to_message": ([dp, dp, dp, dp, dp], None),
    }
    for name in EXPORTS:
        fn = getattr(L, name)      # AttributeError if the symbol is missing
        fn.argtypes, fn.restype = sig[name]
    _LIB = L
    return L
