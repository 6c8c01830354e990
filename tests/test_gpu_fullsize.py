"""Size-independent properties at the sizes BASELINE.json names (the oracle is too slow there): 2048^2 tiles, 4096-wide
phase ROIs, full-frame SURF, 20k x 19k descriptor sets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0
    return g


@pytest.fixture(scope="module")
def pair2048():
    from imagestitch_b200 import synth
    return synth.pair(seed=1234, size=2048, overlap=205, direction=1)        # BASELINE configs[1]


def test_c2_alignment_recovers_true_offset(gpu, pair2048):
    A, B, off = pair2048
    L = int(np.floor(2048 * 0.2))
    r = gpu.align_batch(A[None, 2048 - L:], B[None, :L])[0]
    assert r["status"] == 1 and r["flags"] == 0
    assert abs(int(r["d_row"]) + 2048 - L - off[0]) <= 1 and abs(int(r["d_col"]) - off[1]) <= 1
    assert r["n_a"] == int(0.01 * L * 2048) and r["n_b"] == int(0.01 * L * 2048)      # keypointsRatio cap reached
    # direction-2 strips of the transposed scene give the transposed offset (strided, non-contiguous input)
    At, Bt = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)
    r2 = gpu.align_batch(At[None, :, 2048 - L:], Bt[None, :, :L])[0]
    assert r2["status"] == 1 and abs(int(r2["d_col"]) + 2048 - L - off[0]) <= 1 and abs(int(r2["d_row"]) - off[1]) <= 1


def test_full_frame_surf_properties(gpu, pair2048):
    A, _, _ = pair2048
    kp, d = gpu.surf_detect_and_describe(A, extended=True, keypoints_ratio=0.01)
    assert len(kp) == int(0.01 * 2048 * 2048)                                # 41 943: the plugin's cap
    assert np.all(np.diff(kp[:, 4]) <= 0)                                    # response-ordered (KeypointGreater)
    assert kp[:, 0].min() >= 0 and kp[:, 0].max() < 2048 and kp[:, 1].min() >= 0 and kp[:, 1].max() < 2048
    assert np.all(kp[:, 4] > 100.0) and np.all((kp[:, 3] >= 0) & (kp[:, 3] <= 360))
    assert np.abs(np.linalg.norm(d, axis=1) - 1.0).max() < 1e-4              # unit descriptors
    # translation covariance: an integer crop shifts octave-0 keypoints by exactly the crop origin
    kp2, _ = gpu.surf_detect_and_describe(A[64:1600, 32:1700], extended=True, keypoints_ratio=0.0)
    kpa, _ = gpu.surf_detect_and_describe(A, extended=True, keypoints_ratio=0.0)
    s_full = {(round(float(r[0]) - 32, 3), round(float(r[1]) - 64, 3)) for r in kpa if r[5] == 0}
    inner = [(round(float(r[0]), 3), round(float(r[1]), 3)) for r in kp2 if r[5] == 0 and 80 < r[0] < 1500 and 80 < r[1] < 1400]
    assert len(inner) > 1000 and sum(p in s_full for p in inner) >= 0.99 * len(inner)


def test_matcher_at_full_frame_sizes(gpu):
    """~20k x 19k x 128 (the dendritic ROI sizes of BASELINE.md): tensor-core path == exact SIMT path."""
    rng = np.random.default_rng(0)
    B = rng.standard_normal((19477, 128)).astype(np.float32); B /= np.linalg.norm(B, axis=1, keepdims=True)
    A = rng.standard_normal((20392, 128)).astype(np.float32); A /= np.linalg.norm(A, axis=1, keepdims=True)
    A[:6000] = B[rng.permutation(19477)[:6000]] + 0.04 * rng.standard_normal((6000, 128)).astype(np.float32)
    gpu.set_matcher("tc")
    m_tc = gpu.match_descriptors(A, B, 2, 0.75)
    fb = gpu.last_match_fallbacks()
    gpu.set_matcher("simt")
    m_simt = gpu.match_descriptors(A, B, 2, 0.75)
    gpu.set_matcher("tc")
    assert np.array_equal(m_tc, m_simt) and 5000 < len(m_tc) <= 6100 and fb < 200
    assert np.all(np.diff(m_tc[:, 1]) > 0)                                   # ascending queryIdx


def test_c3_phase_correlation_4096(gpu):
    """BASELINE configs[2]: 819 x 4096 ROI (DFT 864 x 4096): antisymmetry and agreement with cv2."""
    import cv2
    from imagestitch_b200 import synth
    base = synth.canvas(99, 900, 4200)
    a = np.ascontiguousarray(base[10:829, 50:4146]); b = np.ascontiguousarray(base[15:834, 38:4134])    # b = a shifted by (+5, -12)
    (sx, sy), resp = gpu.phase_correlate(a, b)
    (tx, ty), tresp = gpu.phase_correlate(b, a)
    assert abs(sx + tx) < 1e-6 and abs(sy + ty) < 1e-6 and abs(resp - tresp) < 1e-9
    assert abs(sx - 12) < 0.1 and abs(sy + 5) < 0.1 and resp > 0.5
    (cx, cy), cresp = cv2.phaseCorrelate(np.float64(a), np.float64(b))
    assert abs(sx - cx) < 1e-6 and abs(sy - cy) < 1e-6 and abs(resp - cresp) < 1e-9


def test_vote_and_blend_idempotence(gpu):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (205, 2048)).astype(np.int16)
    out1 = gpu.fuse_roi(a, a, "fadeInAndFadeOut", 1843, 2)
    # blending an image with itself: the reference's ramp weights sum to (n+1)/n for dx > 0 (quirk Q4), so the result is
    # trunc(min(255, a * (rows+1)/rows)): never below a, at most 2 levels above
    d = out1.astype(int) - a
    assert out1.shape == a.shape and d.min() >= 0 and d.max() <= 2
    assert np.array_equal(gpu.fuse_roi(a, a, "average"), np.where(a == 0, 0, a).astype(np.uint8))
    assert np.array_equal(gpu.fuse_roi(a, a, "maximum"), gpu.fuse_roi(a, a, "minimum"))


def _oracle_align(roiA, roiB):
    from oracle import surf
    mf = int(0.01 * roiA.size)
    kA, dA = surf.detect_and_compute(roiA, 100, 4, 3, True, False, mf)
    kB, dB = surf.detect_and_compute(roiB, 100, 4, 3, True, False, mf)
    m = surf.match_l2_ratio(dA, dB, 0.75)
    return (kA, dA), (kB, dB), m, surf.offset_by_mode(kA, kB, m, 3)


def _assert_surf_equal(got, want):
    (kg, dg), (ko, do) = got, want
    assert kg.shape == ko.shape and dg.shape == do.shape
    assert np.array_equal(kg[:, :7], ko[:, :7])                              # x, y, size, angle, response, octave, laplacian
    assert np.array_equal(dg, do)                                            # descriptors bit for bit


def test_c2_headline_roi_equals_oracle(gpu, pair2048):
    """BASELINE configs[1] at the size bench.py times (seed 1234, ROI 409 x 2048, GPU-SURF parameters): keypoints, descriptors,
    match list, vote and the batched entry point against the CPU oracle (ImageUtility.py:23-28,272,288-296,139-178)."""
    A, B, off = pair2048
    L = int(np.floor(2048 * 0.2))
    roiA, roiB = np.ascontiguousarray(A[2048 - L:]), np.ascontiguousarray(B[:L])
    fa, fb, m_o, (st_o, off_o, votes_o) = _oracle_align(roiA, roiB)
    ga = gpu.surf_detect_and_describe(roiA, extended=True, keypoints_ratio=0.01)
    gb = gpu.surf_detect_and_describe(roiB, extended=True, keypoints_ratio=0.01)
    _assert_surf_equal(ga, fa); _assert_surf_equal(gb, fb)
    assert np.array_equal(gpu.match_descriptors(ga[1], gb[1], 2, 0.75), m_o)
    r = gpu.align_batch(roiA[None], roiB[None])[0]
    assert (int(r["n_a"]), int(r["n_b"]), int(r["n_matches"])) == (len(fa[0]), len(fb[0]), len(m_o))
    assert (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])], int(r["votes"])) == (st_o, off_o, votes_o)
    assert st_o and abs(off_o[0] + 2048 - L - off[0]) <= 1 and abs(off_o[1] - off[1]) <= 1


def test_c1_iron_roi_pair_equals_oracle(gpu, golden_dir):
    """BASELINE configs[0]: the 387 x 2584 ROI strips of demoImages/iron (committed as PNG fixtures) through align_batch."""
    import os
    import cv2
    roiA = cv2.imread(os.path.join(golden_dir, "iron_A_dir1.png"), cv2.IMREAD_GRAYSCALE)
    roiB = cv2.imread(os.path.join(golden_dir, "iron_B_dir1.png"), cv2.IMREAD_GRAYSCALE)
    fa, fb, m_o, (st_o, off_o, votes_o) = _oracle_align(roiA, roiB)
    _assert_surf_equal(gpu.surf_detect_and_describe(roiA, extended=True, keypoints_ratio=0.01), fa)
    r = gpu.align_batch(roiA[None], roiB[None])[0]
    assert (int(r["n_a"]), int(r["n_b"]), int(r["n_matches"])) == (len(fa[0]), len(fb[0]), len(m_o))
    assert (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])], int(r["votes"])) == (st_o, off_o, votes_o)
    assert st_o and abs(off_o[0] + 1936 - 387 - 1698) <= 1 and off_o[1] == 0           # [1698..1699, 0]
