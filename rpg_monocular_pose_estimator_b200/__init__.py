"""rpg_monocular_pose_estimator_b200 — B200-native hot path of uzh-rpg/rpg_monocular_pose_estimator.

Host-side mirror of the reference's interface for the per-frame path (LEDDetector / PoseEstimator) on top of
the C ABI in include/mpe_b200.h (libmpe_b200.so: hand-written sm_100a CUDA kernels).  There is no CPU
fallback: importing the native library fails loudly when it is missing or when no GPU is present at call
time.
"""
from .synth import Params  # noqa: F401
from ._lib import load_library, MpeError, MpeResult, MPE_MAX_LEDS, MPE_MAX_DET, MPE_MAX_BLOBS  # noqa: F401
from .led_detector import LEDDetector  # noqa: F401
from .pose_estimator import PoseEstimator, Context  # noqa: F401
from .visualization import Visualization  # noqa: F401
