"""GPU: the reference's three-function plugin API (appendix/myGpuFeatures.cpp:67-104, 106-146, 148-195), re-created in
myGpuFeatures/myGpuFeatures.py over libvfsms.so, called on a device exactly as ImageUtility.py:272,274,306,308 call it, unpacked with
the reference's own helper semantics (npToKpsAndDescriptors ImageUtility.py:236-246, npToListForMatches :224-234) and compared
with the CPU oracle (SURF, L2 matcher) and with cv2 (ORB keypoints, Hamming matcher)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plugin():
    from imagestitch_b200 import gpu
    assert gpu.device_count() > 0, "no CUDA device: the plugin has no CPU fallback"
    from myGpuFeatures import myGpuFeatures
    return myGpuFeatures


def _unpack_kps(array):        # ImageUtility.py:236-246, statement for statement in behaviour
    kps = [[array[i, 0, 0], array[i, 1, 0]] for i in range(array.shape[0])]
    return kps, array[:, :, 1]


def _unpack_matches(array):    # ImageUtility.py:224-234
    return [(array[i, 0], array[i, 1]) for i in range(array.shape[0])]


@pytest.mark.parametrize("extended", [True, False])
def test_detect_and_describe_by_surf_layout_and_values(plugin, synth_pair_rois, extended):
    from oracle import surf
    roiA, _, _ = synth_pair_rois
    view = roiA[:, 3:-5]                                           # the callers pass strided ROI views
    arr = plugin.detectAndDescribeBySurf(view, 100.0, 4, 3, extended, 0.01, False)       # ImageUtility.py:272 argument order
    D = 128 if extended else 64
    assert arr.dtype == np.float32 and arr.ndim == 3 and arr.shape[1:] == (D, 2)
    kps, desc = _unpack_kps(arr)
    ko, do = surf.detect_and_compute(np.ascontiguousarray(view), 100.0, 4, 3, extended, False, int(0.01 * view.size))
    assert len(kps) == len(ko) > 300
    assert np.array_equal(np.float32(kps), ko[:, :2])              # (x, y) in plane 0, rows 0 / 1
    assert np.array_equal(desc, do)                                # descriptor in plane 1
    assert not arr[:, 2:, 0].any()                                 # everything else zero (cpp:22-49)


def test_surf_empty_result_is_an_empty_array(plugin):
    arr = plugin.detectAndDescribeBySurf(np.full((64, 96), 90, np.uint8), 100.0, 4, 3, True, 0.01, False)
    assert arr.shape == (0, 128, 2)                                # the plugin returned None here and crashed its caller
    m = plugin.matchDescriptors(np.zeros((0, 128), np.float32), np.zeros((5, 128), np.float32), 2, 0.75)
    assert m.shape == (0, 2) and m.dtype == np.int32


def test_match_descriptors_type2_is_knn2_ratio(plugin, synth_pair_rois):
    from oracle import surf
    roiA, roiB, true_off = synth_pair_rois
    a = plugin.detectAndDescribeBySurf(roiA, 100.0, 4, 3, True, 0.01, False)
    b = plugin.detectAndDescribeBySurf(roiB, 100.0, 4, 3, True, 0.01, False)
    (kA, dA), (kB, dB) = _unpack_kps(a), _unpack_kps(b)
    m = plugin.matchDescriptors(dA, dB, 2, 0.75)                   # ImageUtility.py:306: featureType 2 (surf), param = searchRatio
    assert m.dtype == np.int32 and m.ndim == 2 and m.shape[1] == 2
    assert np.array_equal(m, surf.match_l2_ratio(np.ascontiguousarray(dA), np.ascontiguousarray(dB), 0.75))
    matches = _unpack_matches(m)                                   # rows are (trainIdx, queryIdx), query ascending (cpp:53-65)
    assert all(matches[k][1] < matches[k + 1][1] for k in range(len(matches) - 1))
    # getOffsetByMode on the unpacked lists (ImageUtility.py:139-178): the pair's true offset
    votes = {}
    for t, q in matches:
        key = (int(kA[q][1] - kB[t][1]), int(kA[q][0] - kB[t][0]))
        if key != (0, 0):
            votes[key] = votes.get(key, 0) + 1
    best = max(votes.items(), key=lambda kv: kv[1])
    assert abs(best[0][0] - true_off[0]) <= 1 and abs(best[0][1] - true_off[1]) <= 1 and best[1] >= 3


def test_detect_and_describe_by_orb_and_type3_matcher(plugin, synth_pair_rois):
    import cv2
    roiA, roiB, true_off = synth_pair_rois
    args = (5000, 1.2, 8, 31, 0, 2, 0, 31, 20, True)              # ImageUtility.py:274 argument order (attrs :31-40)
    a = plugin.detectAndDescribeByOrb(roiA, *args)
    b = plugin.detectAndDescribeByOrb(roiB, *args)
    assert a.dtype == np.float32 and a.shape[1:] == (32, 2) and 500 < a.shape[0] <= 5000
    (kA, dA), (kB, dB) = _unpack_kps(a), _unpack_kps(b)
    assert dA.min() >= 0 and dA.max() <= 255 and np.array_equal(dA, np.round(dA))        # descriptor bytes as floats (cpp:118)
    # keypoints: the same detector as cv2.ORB_create with these parameters (tolerances of tests/test_gpu_orb.py)
    ref = cv2.ORB_create(5000, 1.2, 8, 31, 0, 2, 0, 31, 20).detect(np.ascontiguousarray(roiA), None)
    ref_pts = {(round(k.pt[0], 2), round(k.pt[1], 2)) for k in ref if k.octave == 0}
    got = {(round(float(x), 2), round(float(y), 2)) for x, y in kA}
    assert len(ref_pts & got) >= 0.9 * len(ref_pts)
    # matcher type 3: Hamming best-1 with distance < param (cpp:174-187); identical to cv2's BFMatcher on the same bytes
    m = plugin.matchDescriptors(dA, dB, 3, 30)                     # ImageUtility.py:308: param = orbMaxDistance
    bf = cv2.BFMatcher(cv2.NORM_HAMMING).match(np.uint8(dA), np.uint8(dB))
    want = np.int32([(x.trainIdx, x.queryIdx) for x in sorted(bf, key=lambda x: x.queryIdx) if x.distance < 30]).reshape(-1, 2)
    assert np.array_equal(m, want) and len(m) > 50
    votes = {}
    for t, q in _unpack_matches(m):
        key = (int(kA[q][1] - kB[t][1]), int(kA[q][0] - kB[t][0]))
        if key != (0, 0):
            votes[key] = votes.get(key, 0) + 1
    best = max(votes.items(), key=lambda kv: kv[1])
    assert abs(best[0][0] - true_off[0]) <= 1 and abs(best[0][1] - true_off[1]) <= 1
