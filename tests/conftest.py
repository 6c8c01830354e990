import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    ref = os.path.isdir("/root/reference")
    for it in items:
        if "reference" in it.keywords and not ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present on this box"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def synth_pair_rois():
    """ROI strips (i=1, direction 1, roiRatio 0.2) of a seeded 1024^2 synthetic pair + true ROI-relative offset."""
    import numpy as np
    from imagestitch_b200 import synth
    A, B, off = synth.pair(seed=77, size=1024, overlap=110, direction=1)
    L = int(np.floor(1024 * 0.2))
    return A[1024 - L:, :], B[:L, :], (off[0] - (1024 - L), off[1])
