import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "face-landmark-detector_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    with np.load(os.path.join(ROOT, "tests", "golden", "reference_host.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; GPU tests must never pass on a fallback."""
    sys.path.insert(0, PKG)
    import build as fld_build
    fld_build.build()
    from keypoints_detector import _native
    lib = _native.load_library()
    _native.handle()  # raises without an sm_100 device
    return lib
