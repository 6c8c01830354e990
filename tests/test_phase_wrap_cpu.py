"""CPU: the wrap-aware phase-correlation mode (imagestitch_b200/phase_wrap.py) recovers the true offset
of synthetic pairs -- including overlaps below half the ROI, where the plain peak aliases -- with the NumPy oracle standing in
for the two device calls; and the reference-parity mode really is wrong on the same inputs (quirk Q6)."""
import numpy as np
import pytest

from imagestitch_b200 import phase_wrap as pw
from oracle import numpy_oracle as no


def test_optimal_dft_size_matches_cv2():
    import cv2
    for n in (1, 2, 7, 97, 204, 387, 409, 819, 1638, 2457, 2584, 4096):
        assert pw.optimal_dft_size(n) == cv2.getOptimalDFTSize(n) == no.optimal_dft_size(n)


def test_overlap_sums_oracle_against_brute_force():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (9, 13)); b = rng.integers(0, 256, (9, 13))
    shifts = [(0, 0), (3, -4), (-8, 12), (9, 0), (2, 2)]
    got = no.overlap_sums(a, b, shifts)
    for k, (dr, dc) in enumerate(shifts):
        acc = np.zeros(6, np.int64)
        for r in range(9):
            for c in range(13):
                if 0 <= r + dr < 9 and 0 <= c + dc < 13:
                    pa, pb = int(a[r + dr, c + dc]), int(b[r, c])
                    acc += (1, pa, pb, pa * pb, pa * pa, pb * pb)
        assert np.array_equal(got[k], acc)
    assert pw.zncc(no.overlap_sums(a, a, [(0, 0)])[0]) == pytest.approx(1.0)
    assert pw.zncc((0, 0, 0, 0, 0, 0)) == 0.0 and pw.zncc((5, 50, 50, 500, 500, 500)) == 0.0      # empty / flat


@pytest.mark.parametrize("direction", [1, 2])
@pytest.mark.parametrize("overlap", [150, 100, 60, 30])
def test_wrap_aware_recovers_true_offset(direction, overlap):
    from imagestitch_b200 import synth
    from imagestitch_b200.ImageUtility import Method
    size, ratio = 512, 0.4                       # ROI strip 204 px: overlaps 100, 60, 30 are below half of it -> the peak aliases
    A, B, off = synth.pair(seed=100 + overlap + direction, size=size, overlap=overlap, direction=direction)
    m = Method()
    roiA = m.getROIRegionForIncreMethod(A, direction=direction, order="first", searchRatio=ratio)
    roiB = m.getROIRegionForIncreMethod(B, direction=direction, order="second", searchRatio=ratio)
    st, d, score, resp = pw.resolve(roiA, roiB, no.phase_correlate, no.overlap_sums)
    L = int(size * ratio)
    full = [d[0] + (size - L if direction == 1 else 0), d[1] + (size - L if direction == 2 else 0)]
    assert st and score > 0.8, (st, score)
    assert abs(full[0] - off[0]) <= 1 and abs(full[1] - off[1]) <= 1, (full, off)
    # the parity mode on the same ROIs: the reference adds (int(sy), int(sx)) -> not the true offset
    (sx, sy), _ = no.phase_correlate(roiA, roiB)
    ref_full = [int(sy) + (size - L if direction == 1 else 0), int(sx) + (size - L if direction == 2 else 0)]
    assert abs(ref_full[0] - off[0]) > 1 or abs(ref_full[1] - off[1]) > 1


def test_wrap_aware_rejects_unrelated_tiles():
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=1, size=256, overlap=60, direction=1)
    C, _, _ = synth.pair(seed=2, size=256, overlap=60, direction=1)
    st, d, score, resp = pw.resolve(A[-100:], C[:100], no.phase_correlate, no.overlap_sums)
    assert not st and score < 0.5


def test_stitcher_wrap_aware_mode_end_to_end_on_cpu(monkeypatch):
    """Stitcher.calculateOffsetForPhaseCorrleateIncre with phaseMode = "wrapAware": the full tile offset (ROI origin added back
    by the shared search loop) is the true one; the two device calls are replaced by the NumPy oracle."""
    from imagestitch_b200 import gpu, synth
    from imagestitch_b200.Stitcher import Stitcher
    monkeypatch.setattr(gpu, "phase_correlate", lambda a, b, device=0: no.phase_correlate(a, b))
    monkeypatch.setattr(gpu, "overlap_sums", lambda a, b, s, device=0: no.overlap_sums(a, b, s))
    monkeypatch.setattr(Stitcher, "isPrintLog", False)
    monkeypatch.setattr(Stitcher, "phaseMode", "wrapAware")
    monkeypatch.setattr(Stitcher, "roiRatio", 0.2)
    monkeypatch.setattr(Stitcher, "direction", 1)
    monkeypatch.setattr(Stitcher, "directIncre", 1)
    st = Stitcher()
    for seed, direction, overlap in ((5, 1, 40), (6, 2, 70), (7, 1, 90)):
        A, B, off = synth.pair(seed=seed, size=512, overlap=overlap, direction=direction)
        status, offset = st.calculateOffsetForPhaseCorrleateIncre([A, B])
        assert status and abs(offset[0] - off[0]) <= 1 and abs(offset[1] - off[1]) <= 1, (offset, off)
        assert st.direction == direction
    if "direction" in st.__dict__:
        del st.__dict__["direction"]
