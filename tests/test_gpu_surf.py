"""GPU parity: SURF / matcher / vote through the C ABI vs the CPU oracle on the same inputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load_png(path):
    import cv2
    return cv2.imread(path, cv2.IMREAD_GRAYSCALE)


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return g


def _compare_surf(gpu, img, extended, ratio, upright=False, thr=100.0):
    from oracle import surf
    h, w = img.shape
    mf = int(min(max(ratio * h * w, 1), 65535)) if ratio > 0 else 0
    kp_o, d_o = surf.detect_and_compute(img, thr, 4, 3, extended, upright, mf)
    kp_g, d_g = gpu.surf_detect_and_describe(img, hessian_threshold=thr, extended=extended, keypoints_ratio=ratio,
                                             upright=upright)
    assert len(kp_g) == len(kp_o), (len(kp_g), len(kp_o))
    # keypoint geometry and order: bit-exact (x, y, size, response, octave, laplacian)
    for col in (0, 1, 2, 4, 5, 6):
        assert np.array_equal(kp_g[:, col], kp_o[:, col]), "column %d differs" % col
    # orientation: same polynomial atan, same summation order -> exact
    assert np.array_equal(kp_g[:, 3], kp_o[:, 3])
    # descriptors: bit for bit (the oracle takes the correctly rounded sin / cos of the orientation, like the device code)
    assert np.array_equal(d_g, d_o), (np.abs(d_g - d_o).max(), int((np.abs(d_g - d_o).max(axis=1) > 0).sum()))
    return kp_g, d_g, kp_o, d_o


def test_surf_synthetic_roi_128(gpu, synth_pair_rois):
    roiA, roiB, _ = synth_pair_rois
    _compare_surf(gpu, roiA, True, 0.01)


def test_surf_synthetic_roi_64_unlimited(gpu, synth_pair_rois):
    roiA, roiB, _ = synth_pair_rois
    _compare_surf(gpu, roiB, False, 0.0)


def test_surf_upright(gpu, synth_pair_rois):
    roiA, _, _ = synth_pair_rois
    _compare_surf(gpu, roiA, True, 0.01, upright=True)


def test_surf_strided_view(gpu, synth_pair_rois):
    """direction-2 ROIs are non-contiguous column strips (ImageUtility.py:84-89)."""
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=5, size=512, overlap=60, direction=2)
    view = A[:, 512 - 102:]
    assert not view.flags["C_CONTIGUOUS"]
    _compare_surf(gpu, view, True, 0.01)


def test_surf_real_micrograph_roi(gpu, golden_dir):
    img = _load_png(os.path.join(golden_dir, "iron_A_dir1.png"))
    _compare_surf(gpu, img, True, 0.01)


def test_surf_empty_and_tiny(gpu):
    flat = np.full((64, 96), 128, np.uint8)
    kp, d = gpu.surf_detect_and_describe(flat)
    assert kp.shape == (0, 8) and d.shape == (0, 128)
    tiny = np.random.default_rng(0).integers(0, 255, (12, 300), dtype=np.uint8)   # smaller than the 2nd layer
    kp, d = gpu.surf_detect_and_describe(tiny)
    from oracle import surf
    kp_o, _ = surf.detect_and_compute(tiny, 100, 4, 3, True, False, 36)
    assert len(kp) == len(kp_o)


def test_match_and_vote_parity(gpu, synth_pair_rois):
    from oracle import surf
    roiA, roiB, true_off = synth_pair_rois
    kA, dA = gpu.surf_detect_and_describe(roiA)
    kB, dB = gpu.surf_detect_and_describe(roiB)
    m_g = gpu.match_descriptors(dA, dB, 2, 0.75)
    m_o = surf.match_l2_ratio(dA, dB, 0.75)
    assert np.array_equal(m_g, m_o)
    st_g, off_g, votes_g = gpu.offset_by_mode(kA, kB, m_g, 3)
    st_o, off_o, votes_o = surf.offset_by_mode(kA, kB, m_o, 3)
    assert (st_g, off_g, votes_g) == (st_o, off_o, votes_o)
    assert st_g and abs(off_g[0] - true_off[0]) <= 1 and abs(off_g[1] - true_off[1]) <= 1


def test_match_edge_cases(gpu):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((70, 64)).astype(np.float32)
    B = rng.standard_normal((1, 64)).astype(np.float32)
    assert gpu.match_descriptors(A, B, 2, 0.75).shape == (0, 2)          # one train row: no second neighbour
    assert gpu.match_descriptors(A[:0], B, 2, 0.75).shape == (0, 2)
    # duplicated train rows: ties resolve to the lower train index, ratio test then fails (d0 == d1)
    B2 = np.repeat(A[:5], 2, axis=0)
    from oracle import surf
    assert np.array_equal(gpu.match_descriptors(A, B2, 2, 0.75), surf.match_l2_ratio(A, B2, 0.75))
    A3 = rng.standard_normal((300, 128)).astype(np.float32); B3 = rng.standard_normal((257, 128)).astype(np.float32)
    assert np.array_equal(gpu.match_descriptors(A3, B3, 2, 0.9), surf.match_l2_ratio(A3, B3, 0.9))


def test_vote_tie_and_zero_rules(gpu):
    # two offsets with equal counts: the first seen in match order wins; exact (0,0) is dropped
    kA = np.array([[10, 10], [20, 20], [30, 30], [40, 40], [5, 5]], np.float32)
    kB = np.array([[8, 7], [18, 17], [29, 26], [39, 36], [5, 5]], np.float32)
    m = np.array([[2, 2], [0, 0], [3, 3], [1, 1], [4, 4]], np.int32)     # offsets (4,1),(3,2),(4,1),(3,2),(0,0)
    st, off, votes = gpu.offset_by_mode(kA, kB, m, 2)
    assert (st, off, votes) == (True, [4, 1], 2)
    st, off, votes = gpu.offset_by_mode(kA, kB, m[[4]], 1)
    assert (st, off, votes) == (True, [0, 0], 1)
    st, off, votes = gpu.offset_by_mode(kA, kB, m[:0], 1)
    assert (st, off, votes) == (False, [0, 0], 0)
    # truncation toward zero: int(-1.9) = -1
    kA2 = np.array([[0.0, 0.1]] * 3, np.float32); kB2 = np.array([[0.0, 2.0]] * 3, np.float32)
    st, off, votes = gpu.offset_by_mode(kA2, kB2, np.array([[0, 0], [1, 1], [2, 2]], np.int32), 3)
    assert (st, off, votes) == (True, [-1, 0], 3)


def test_align_batch_matches_stagewise(gpu, synth_pair_rois):
    from oracle import surf
    roiA, roiB, true_off = synth_pair_rois
    res = gpu.align_batch(np.stack([roiA, roiA]), np.stack([roiB, roiB]))
    kA, dA = surf.detect_and_compute(roiA, 100, 4, 3, True, False, int(0.01 * roiA.size))
    kB, dB = surf.detect_and_compute(roiB, 100, 4, 3, True, False, int(0.01 * roiB.size))
    m = surf.match_l2_ratio(dA, dB, 0.75)
    st, off, votes = surf.offset_by_mode(kA, kB, m, 3)
    for r in res:
        assert r["n_a"] == len(kA) and r["n_b"] == len(kB)
        assert int(r["n_matches"]) == len(m)
        assert bool(r["status"]) == st and [int(r["d_row"]), int(r["d_col"])] == off
        assert int(r["votes"]) == votes
    assert abs(off[0] - true_off[0]) <= 1 and abs(off[1] - true_off[1]) <= 1


def test_align_batches_stream_equals_one_call_per_batch(gpu):
    """The double-buffered form (vfsms_align_batch_upload / _run): batches of different content, pair count and ROI
    shape, pinned and pageable host memory -- every batch's results equal gpu.align_batch on that batch."""
    import torch
    from imagestitch_b200 import synth
    batches = []
    for k, (P, size, overlap) in enumerate([(3, 384, 96), (2, 384, 96), (1, 320, 80), (4, 384, 128), (3, 384, 96)]):
        A = np.empty((P, overlap, size), np.uint8); B = np.empty_like(A)
        for p in range(P):
            a, b, _ = synth.pair(seed=100 + 10 * k + p, size=size, overlap=overlap, direction=1)
            A[p] = a[size - overlap:]; B[p] = b[:overlap]
        if k % 2 == 0 and torch.cuda.is_available():      # pinned: the upload really is asynchronous
            tA = torch.from_numpy(A).pin_memory(); tB = torch.from_numpy(B).pin_memory()
            A, B = tA.numpy(), tB.numpy()
            batches.append((A, B, tA, tB))
        else:
            batches.append((A, B))
    assert list(gpu.align_batches(iter(()))) == []                       # nothing in, nothing out
    one = list(gpu.align_batches([(batches[1][0], batches[1][1])]))      # a single batch: upload, run
    assert len(one) == 1 and np.array_equal(one[0], gpu.align_batch(batches[1][0], batches[1][1]))
    streamed = list(gpu.align_batches((b[0], b[1]) for b in batches))
    assert len(streamed) == len(batches)
    for b, got in zip(batches, streamed):
        want = gpu.align_batch(b[0], b[1])
        assert np.array_equal(got, want)
    # a run without an upload is an error, not stale data
    from imagestitch_b200 import _lib
    import ctypes
    res = np.zeros(1, gpu.PAIR_RESULT_DTYPE)
    p = gpu.surf_params()
    rc = _lib.load().vfsms_align_batch_run(_lib.context(0), 0, ctypes.byref(p), 0.75, 3, res.ctypes.data_as(ctypes.c_void_p))
    assert rc != 0 and b"no uploaded batch" in _lib.load().vfsms_last_error()
