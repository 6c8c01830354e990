"""GPU: the overlap-sum reduction (vfsms_overlap_sums_host) is integer-exact against the NumPy oracle, and the wrap-aware
phase mode recovers true offsets through the device calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return g


def test_overlap_sums_exact(gpu):
    from oracle import numpy_oracle as no
    rng = np.random.default_rng(0)
    for (rows, cols) in ((9, 13), (204, 512), (409, 2048), (1, 1)):
        a = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        b = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        shifts = [(0, 0), (rows // 3, -(cols // 4)), (-(rows - 1), cols - 1), (rows, 0), (0, -cols), (1, 1)]
        assert np.array_equal(gpu.overlap_sums(a, b, shifts), no.overlap_sums(a, b, shifts)), (rows, cols)
    big = np.full((2048, 2048), 255, np.uint8)                       # largest sums: 255^2 * 4 Mpx
    assert np.array_equal(gpu.overlap_sums(big, big, [(0, 0)]), no.overlap_sums(big, big, [(0, 0)]))
    view = np.asfortranarray(rng.integers(0, 256, (64, 80)).astype(np.uint8))[:, 8:72]     # strided column strip (direction 2 / 4)
    assert np.array_equal(gpu.overlap_sums(view, view, [(3, 5)]), no.overlap_sums(view, view, [(3, 5)]))


def test_wrap_aware_mode_recovers_true_offsets(gpu):
    from imagestitch_b200 import synth
    from imagestitch_b200.Stitcher import Stitcher
    st = Stitcher()
    Stitcher.isPrintLog = False
    try:
        Stitcher.phaseMode = "wrapAware"; Stitcher.roiRatio = 0.2; Stitcher.direction = 1; Stitcher.directIncre = 1
        for seed, direction, overlap in ((5, 1, 40), (6, 2, 70), (7, 1, 90), (8, 2, 50)):
            A, B, off = synth.pair(seed=seed, size=512, overlap=overlap, direction=direction)
            status, offset = st.calculateOffsetForPhaseCorrleateIncre([A, B])
            assert status and abs(offset[0] - off[0]) <= 1 and abs(offset[1] - off[1]) <= 1, (offset, off)
    finally:
        Stitcher.phaseMode = "reference"; Stitcher.direction = 1; Stitcher.isPrintLog = True
        if "direction" in st.__dict__:
            del st.__dict__["direction"]
