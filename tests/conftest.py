import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# VFSMS_EMU=1 (set only by tests/test_kernels_emulated_cpu.py for its pytest subprocess): the `gpu` tests of this directory
# run against tests/cuda_emu/_build/libvfsms_emu.so -- the same .cu sources compiled by g++ on a CUDA-on-CPU execution model.
# Test infrastructure only: the product (imagestitch_b200._lib) knows nothing about it and still fails without libvfsms.so.
EMU = os.environ.get("VFSMS_EMU") == "1"
if EMU:
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
    import build_emu
    from imagestitch_b200 import _lib as _vfsms_lib
    _vfsms_lib.SO_PATH = build_emu.build()
    _real_context = _vfsms_lib.context

    def _emu_context(device=0, lane=0):
        fresh = (device if lane == 0 else (device, lane)) not in _vfsms_lib._contexts
        h = _real_context(device, lane)
        if fresh:       # match_tc.cu (tcgen05 inline PTX) is not emulated: the exact SIMT matcher, whatever a test selects
            L = _vfsms_lib.load()
            set_matcher = L.vfsms_set_matcher
            if not getattr(set_matcher, "_emu", False):
                wrapper = lambda ctx, mode: set_matcher(ctx, 1)
                wrapper._emu = True
                L.vfsms_set_matcher = wrapper
            set_matcher(h, 1)
        return h
    _vfsms_lib.context = _emu_context


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


# under emulation: no tcgen05 / TMA (inline PTX), no torch CUDA tensors, and the full-size cases would take hours
_EMU_SKIP = ("test_gpu_match_tc.py", "test_gpu_fullsize.py", "test_decode_into_device_tile_stack", "test_describe_stacked_texture_row_limit_groups",
             "test_two_contexts_run_a_round_side_by_side")      # the emulation runs one launch at a time: no second host thread


# gpu tests of code that has not run on the B200 yet would be listed here (node-id substrings): they are ordered after everything
# that already passed there, so that under `-x` a first-hardware-run failure cannot hide the proven tests.  Empty: every gpu test
# of the suite has passed on hardware (profiles/r02/r02p_gpu_tests.log).
_AFTER_PROVEN = ()


def pytest_collection_modifyitems(config, items):
    ref = os.path.isdir("/root/reference")
    items.sort(key=lambda it: any(s in it.nodeid for s in _AFTER_PROVEN))       # stable: order inside each group is kept
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for it in items:
        # on hardware: a kernel that never returns blocks inside a C call where no signal handler runs -> the watchdog thread ends the
        # process (and with it the CUDA context) instead of leaving the box hung until the caller's own limit
        if has_timeout and not EMU and "gpu" in it.keywords and it.get_closest_marker("timeout") is None:
            it.add_marker(pytest.mark.timeout(900, method="thread"))
        if EMU and any(s in it.nodeid for s in _EMU_SKIP):
            it.add_marker(pytest.mark.skip(reason="not covered by the CUDA-on-CPU emulation"))
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
