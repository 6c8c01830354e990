"""GPU parity: phase correlation vs cv2.phaseCorrelate (the call the reference makes, Stitcher.py:230) and the
reference-level (status, offset) of calculateOffsetForPhaseCorrleateIncre semantics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0
    return g


def _check(gpu, a, b):
    import cv2
    from oracle import numpy_oracle as no
    (sx, sy), resp = gpu.phase_correlate(a, b)
    (cx, cy), cresp = cv2.phaseCorrelate(np.float64(a), np.float64(b))
    (nx, ny), nresp = no.phase_correlate(a, b)
    assert abs(sx - cx) < 1e-6 and abs(sy - cy) < 1e-6 and abs(resp - cresp) < 1e-9, ((sx, sy, resp), (cx, cy, cresp))
    assert abs(sx - nx) < 1e-6 and abs(sy - ny) < 1e-6 and abs(resp - nresp) < 1e-9
    assert [int(sy), int(sx)] == [int(cy), int(cx)]          # the truncation the reference applies (Stitcher.py:231-232)
    return (sx, sy), resp


def test_phase_synthetic_roi(gpu, synth_pair_rois):
    roiA, roiB, _ = synth_pair_rois
    _check(gpu, roiA, roiB)


def test_phase_non_power_of_two_and_strided(gpu):
    from imagestitch_b200 import synth
    A, B, _ = synth.pair(seed=3, size=600, overlap=90, direction=2)
    a = A[:, 600 - 120:]; b = B[:, :120]           # 600 x 120 -> DFT 600 x 120, strided views
    assert not a.flags["C_CONTIGUOUS"]
    _check(gpu, a, b)
    _check(gpu, A[:387, :517], B[:387, :517])     # 387 -> 400, 517 -> 540


def test_phase_known_shift(gpu):
    rng = np.random.default_rng(0)
    base = (rng.random((300, 420)) * 255).astype(np.uint8)
    import cv2
    base = cv2.GaussianBlur(base, (0, 0), 2.0)
    a = base[20:276, 30:286]; b = base[27:283, 19:275]       # b is a shifted by (+7 rows, -11 cols)
    (sx, sy), resp = _check(gpu, a, b)
    assert abs(sx - 11) < 0.2 and abs(sy + 7) < 0.2 and resp > 0.3


def test_phase_flat_and_tiny(gpu):
    flat = np.full((16, 16), 50, np.uint8)
    (sx, sy), resp = gpu.phase_correlate(flat, flat)
    assert np.isfinite(sx) and np.isfinite(sy) and np.isfinite(resp)
    t = np.arange(12, dtype=np.uint8).reshape(3, 4) * 9
    _check(gpu, t, t[::-1].copy())
