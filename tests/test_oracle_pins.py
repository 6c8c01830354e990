"""Pins of the CPU oracle (oracle/) -- run everywhere, no GPU.

What the reference itself offers as ground truth for this path: the 89 golden offsets at Stitcher.py:87 (dendriticCrystal)
and its demo images; nothing else (it has no tests).  cv2 in this image is the second source: it pins every piece of the
SURF restatement that cv2 can still compute (integral, INTER_AREA resize, fastAtan2, odd Gaussian kernels), the matcher
and phase correlation."""
import json
import os

import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import surf


def test_integral_is_cumsum():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (123, 257), dtype=np.uint8)
    ref = np.zeros((124, 258), np.int64)
    ref[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
    assert np.array_equal(surf.integral(img), ref)
    assert np.array_equal(surf.integral(img[:, 5:100]), ref[:, 100:101] * 0 + np.pad(img[:, 5:100].astype(np.int64).cumsum(0).cumsum(1), ((1, 0), (1, 0))))


def test_inter_area_patch_equals_cv2():
    import cv2
    rng = np.random.default_rng(1)
    for n in list(range(22, 90)) + [105, 126, 147, 168, 211, 333, 420, 739]:
        w = rng.integers(0, 256, (n, n), dtype=np.uint8)
        assert np.array_equal(surf.resize_area_u8(w), cv2.resize(w, (21, 21), interpolation=cv2.INTER_AREA)), n


def test_fast_atan2_equals_cv2():
    import cv2
    rng = np.random.default_rng(2)
    ys = rng.standard_normal(3000).astype(np.float32); xs = rng.standard_normal(3000).astype(np.float32)
    ys[:4] = [0, 1, -1, 0]; xs[:4] = [1, 0, 0, -1]
    for y, x in zip(ys, xs):
        assert surf.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_gaussian_kernel_formula():
    import cv2
    # odd sizes: cv2 4.x still evaluates the plain formula; (even sizes differ in 4.x's bit-exact path, the reference pins 3.3.1)
    assert np.abs(surf.gaussian_kernel(13, 2.5) - cv2.getGaussianKernel(13, 2.5, cv2.CV_32F).ravel()).max() < 2e-8
    x = np.arange(20) - 9.5
    t = np.exp(-0.5 * x * x / 3.3 ** 2).astype(np.float32).astype(np.float64)
    assert np.abs(surf.gaussian_kernel(20, 3.3) - (t / t.sum())).max() < 1e-7


def test_box_filter_responses_against_brute_force():
    """det/trace of one layer vs direct pixel sums (known-answer property of the Hessian boxes)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (40, 48), dtype=np.uint8)
    s = surf.integral(img)
    det, tr = surf.layer_det_trace(s, 9, 1)
    f = img.astype(np.float64)
    i, j = 7, 11                                     # sample -> layer position (i+4, j+4)
    dxx = f[i + 2:i + 7, j:j + 3].sum() - 2 * f[i + 2:i + 7, j + 3:j + 6].sum() + f[i + 2:i + 7, j + 6:j + 9].sum()
    dyy = f[i:i + 3, j + 2:j + 7].sum() - 2 * f[i + 3:i + 6, j + 2:j + 7].sum() + f[i + 6:i + 9, j + 2:j + 7].sum()
    dxy = f[i + 1:i + 4, j + 1:j + 4].sum() - f[i + 1:i + 4, j + 5:j + 8].sum() - f[i + 5:i + 8, j + 1:j + 4].sum() + f[i + 5:i + 8, j + 5:j + 8].sum()
    dxx /= 15.0; dyy /= 15.0; dxy /= 9.0
    assert abs(det[i + 4, j + 4] - (dxx * dyy - 0.81 * dxy * dxy)) < 1e-2 * max(1.0, abs(det[i + 4, j + 4]))
    assert abs(tr[i + 4, j + 4] - (dxx + dyy)) < 1e-3


def test_surf_translation_covariance():
    """Shifting the image content by an integer vector shifts interior octave-0 keypoints by the same vector
    (higher octaves sample on a grid anchored at the image origin, so only even / multiple-of-step shifts carry over)."""
    from imagestitch_b200 import synth
    base = synth.canvas(5, 420, 520)
    a = base[20:320, 30:430]; b = base[27:327, 19:419]          # b(y, x) = a(y + 7, x - 11)
    ka, _ = surf.detect_and_compute(a, 400, 3, 3, False, False, 0, want_desc=False)
    kb, _ = surf.detect_and_compute(b, 400, 3, 3, False, False, 0, want_desc=False)
    sa = {(round(float(r[0]), 3), round(float(r[1]), 3)) for r in ka if 80 < r[0] < 300 and 80 < r[1] < 200 and r[5] == 0}
    sb = {(round(float(r[0]) - 11, 3), round(float(r[1]) + 7, 3)) for r in kb if r[5] == 0}
    assert len(sa) > 20 and len(sa - sb) <= len(sa) // 20


def test_golden_offsets_file(golden_dir):
    """Offset-level pin: reference Stitcher + oracle SURF vs the author's list (Stitcher.py:87); 89/89 within +-1 px."""
    d = json.load(open(os.path.join(golden_dir, "dendritic_offsets.json")))
    gold = d["golden_Stitcher_py_87"]; ours = d["oracle_surf_offsets"]
    assert len(gold) == 89 and len(ours) == 89
    within = sum(1 for o, g in zip(ours, gold) if o[0] and abs(o[1][0] - g[0]) <= 1 and abs(o[1][1] - g[1]) <= 1)
    assert within == d["within1"] == 89 and d["exact"] >= 80


def test_golden_roi_fixture_reproduces_offset(golden_dir):
    """Oracle pipeline on the committed ROI strips of dendritic pair 0 reproduces the golden offset of that pair."""
    import cv2
    d = json.load(open(os.path.join(golden_dir, "dendritic_offsets.json")))
    A = cv2.imread(os.path.join(golden_dir, "dendritic_00_A_dir1.png"), 0); B = cv2.imread(os.path.join(golden_dir, "dendritic_00_B_dir1.png"), 0)
    kA, dA = surf.detect_and_compute(A); kB, dB = surf.detect_and_compute(B)
    st, off, votes = surf.offset_by_mode(kA, kB, surf.match_l2_ratio(dA, dB, 0.75), 3)
    H = d["shape"][0]
    assert st and abs(off[0] + H - int(0.2 * H) - d["golden_Stitcher_py_87"][0][0]) <= 1 and abs(off[1] - d["golden_Stitcher_py_87"][0][1]) <= 1


def test_matcher_oracle_equals_cv2_bfmatcher(golden_dir):
    import cv2
    img = cv2.imread(os.path.join(golden_dir, "iron_A_dir1.png"), 0)[:, :900]
    img2 = cv2.imread(os.path.join(golden_dir, "iron_B_dir1.png"), 0)[:, :900]
    _, dA = surf.detect_and_compute(img, extended=True, max_features=3000); _, dB = surf.detect_and_compute(img2, extended=True, max_features=3000)
    raw = cv2.DescriptorMatcher_create("BruteForce").knnMatch(dA, dB, 2)
    ref = [(m[0].trainIdx, m[0].queryIdx) for m in raw if len(m) == 2 and m[0].distance < m[1].distance * 0.75]
    ours = [tuple(r) for r in surf.match_l2_ratio(dA, dB, 0.75)]
    assert ours == ref


def test_vote_oracle_equals_reference_semantics():
    """Pure-Python transcription of getOffsetByMode (ImageUtility.py:139-178) on random matches."""
    rng = np.random.default_rng(4)
    kA = (rng.random((200, 2)) * 300).astype(np.float32); kB = (rng.random((180, 2)) * 300).astype(np.float32)
    kB[:60] = kA[:60] - np.float32([3.4, 7.7])
    m = np.stack([rng.permutation(180)[:150], rng.permutation(200)[:150]], 1).astype(np.int32)
    m[:60] = np.stack([np.arange(60), np.arange(60)], 1)
    dx, dy = [], []
    for t, q in m:
        r, c = int(kA[q][1] - kB[t][1]), int(kA[q][0] - kB[t][0])
        if r == 0 and c == 0:
            continue
        dx.append(r); dy.append(c)
    z = list(zip(dx, dy))
    cnt = {a: z.count(a) for a in z}
    best = sorted(cnt.items(), key=lambda x: x[1], reverse=True)[0]
    st, off, votes = surf.offset_by_mode(kA, kB, m, 3)
    assert (st, off, votes) == (best[1] >= 3, [best[0][0], best[0][1]], best[1])


def test_phase_oracle_equals_cv2():
    import cv2
    from imagestitch_b200 import synth
    A, B, _ = synth.pair(seed=77, size=512, overlap=60, direction=1)
    for a, b in ((A[512 - 102:], B[:102]), (A[:387, :300], B[:387, 17:317])):
        (cx, cy), cr = cv2.phaseCorrelate(np.float64(a), np.float64(b))
        (nx, ny), nr = no.phase_correlate(a, b)
        assert abs(cx - nx) < 1e-9 and abs(cy - ny) < 1e-9 and abs(cr - nr) < 1e-12
    for n in (1, 7, 387, 409, 819, 1638, 2457, 2584, 4096):
        assert no.optimal_dft_size(n) == cv2.getOptimalDFTSize(n)


@pytest.mark.reference
def test_roi_and_vote_against_unmodified_reference():
    """Build container only: our host-side ROI logic equals the reference's on all directions / orders."""
    from oracle import reference_shims as rs
    from imagestitch_b200.ImageUtility import Method
    S, U, F = rs.import_reference()
    ref, ours = U.Method(), Method()
    img = np.arange(60 * 70, dtype=np.uint8).reshape(60, 70)
    for d in (1, 2, 3, 4):
        for order in ("first", "second"):
            for ratio in (0.2, 0.4, 0.6000000000000001):
                assert np.array_equal(ref.getROIRegionForIncreMethod(img, d, order, ratio), ours.getROIRegionForIncreMethod(img, d, order, ratio))


def test_cuda_path_reproduced_the_golden_vector_on_hardware():
    """profiles/r02/dendritic_89_pairs_gpu.json is the record of scripts/golden_grid_gpu.py on a B200 (the 90 demo JPEGs cannot be
    committed): the CUDA path -- library JPEG decode, batched incremental SURF search -- against the author's list at
    Stitcher.py:87.  The record must agree with the committed golden list pair by pair."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "r02", "dendritic_89_pairs_gpu.json")))
    golden = json.load(open(os.path.join(root, "tests", "golden", "dendritic_offsets.json")))["golden_Stitcher_py_87"]
    assert rec["tiles"] == 90 and rec["shape"] == [1936, 2584]
    for name, run in rec["runs"].items():
        assert run["pairs"] == 89 and len(run["offsets"]) == 89
        ok = sum(1 for (st, off, _), g in zip(run["offsets"], golden) if st and max(abs(off[0] - g[0]), abs(off[1] - g[1])) <= 1)
        assert ok == run["within_1px_of_golden"] == 89, (name, ok)
