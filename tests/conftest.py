import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pose_oracle
    return pose_oracle.lib()


@pytest.fixture(scope="session")
def gpu_ctx_752():
    """One context for 752x480 single-frame stage calls and small batches."""
    import rpg_monocular_pose_estimator_b200 as mpe
    ctx = mpe.Context(0, 64, 752, 480)
    yield ctx
    ctx.close()
