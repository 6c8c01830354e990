"""GPU: every kernel schedule selectable through vfsms_set_option (include/vfsms.h VFSMS_OPT_*) must give results IDENTICAL
to every other one.  The defaults (describe=1: fixed-point chunked sampler, sort=1, lpt=1) are what bench.py times and what the
oracle tests of the rest of the suite run on; describe=0 is the reference sampler (double precision, u8 image)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0
    defaults = {name: g.get_option(name) for name in g.OPTIONS}
    yield g
    for name, v in defaults.items():
        g.set_option(name, v)


def _surf_both(gpu, img, option, values, **kw):
    outs = []
    for v in values:
        gpu.set_option(option, v)
        assert gpu.get_option(option) == v
        outs.append(gpu.surf_detect_and_describe(img, **kw))
    return outs


@pytest.mark.parametrize("extended", [True, False])
def test_describe_samplers_identical(gpu, synth_pair_rois, extended):
    roiA, roiB, _ = synth_pair_rois
    for img in (roiA, roiB):
        outs = _surf_both(gpu, img, "describe", (1, 0), extended=extended, keypoints_ratio=0.01)
        k1, d1 = outs[0]
        assert len(k1) > 500
        for k2, d2 in outs[1:]:
            assert np.array_equal(k1, k2) and np.array_equal(d1, d2)


def test_describe_samplers_borders_and_giants(gpu):
    """small image: almost every window crosses the border; low threshold + 4 octaves: windows up to several hundred px"""
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=3, size=384, overlap=60, direction=1)
    for img in (A[:97], A[:, :131], A):
        outs = _surf_both(gpu, img, "describe", (1, 0), extended=True, keypoints_ratio=0.0, hessian_threshold=30.0)
        (k1, d1), (k2, d2) = outs[0], outs[1]
        assert len(k1) > 50
        assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
        gpu.set_option("describe", 1)
        gpu.surf_detect_and_describe(img, extended=True, keypoints_ratio=0.0, hessian_threshold=30.0)
        assert gpu.last_describe_handovers() < 0.05 * len(k1) + 4          # the fixed-point sampler does the work, not the hand-over


def test_describe_stacked_texture_batches(gpu):
    """align_batch over many ROIs: image index -> texture row offset.  (The CPU emulation runs this test a second time with a
    450-row texture limit, which splits these 12 images into groups like the 80 x 1020-row case below.)"""
    from imagestitch_b200 import synth
    rois_a, rois_b = [], []
    for k in range(6):
        A, B, _ = synth.pair(seed=40 + k, size=512, overlap=64, direction=1)
        rois_a.append(A[512 - 102:]); rois_b.append(B[:102])
    ra, rb = np.stack(rois_a), np.stack(rois_b)
    gpu.set_option("describe", 0); r1 = gpu.align_batch(ra, rb)
    gpu.set_option("describe", 1); r2 = gpu.align_batch(ra, rb)
    assert np.array_equal(r1, r2) and r1["status"].all()


def test_describe_stacked_texture_row_limit_groups(gpu):
    """80 ROIs x 1020 rows exceed the 65000-row limit of one 2-D linear texture: the launch is split into groups."""
    from imagestitch_b200 import synth
    A, B, _ = synth.pair(seed=77, size=1024, overlap=110, direction=2)
    ta = np.stack([np.ascontiguousarray(np.roll(A, 13 * k, axis=0)[:, 1024 - 204:]) for k in range(40)])
    tb = np.stack([np.ascontiguousarray(np.roll(B, 13 * k, axis=0)[:, :204]) for k in range(40)])
    ta = np.ascontiguousarray(ta.transpose(0, 2, 1)); tb = np.ascontiguousarray(tb.transpose(0, 2, 1))   # 204 x 1024 strips
    big_a = np.ascontiguousarray(np.repeat(ta, 5, axis=1)[:, :1024]); big_b = np.ascontiguousarray(np.repeat(tb, 5, axis=1)[:, :1024])
    gpu.set_option("describe", 0); r1 = gpu.align_batch(big_a, big_b)
    gpu.set_option("describe", 1); r2 = gpu.align_batch(big_a, big_b)          # 80 images x 1024 rows -> 2 groups
    assert np.array_equal(r1, r2)


def test_sort_per_image_identical(gpu, synth_pair_rois):
    roiA, roiB, _ = synth_pair_rois
    for kw in (dict(extended=True, keypoints_ratio=0.01), dict(extended=False, keypoints_ratio=0.0),
               dict(extended=True, keypoints_ratio=0.002)):
        (k0, d0), (k1, d1) = _surf_both(gpu, roiA, "sort", (0, 1), **kw)
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1)
    flat = np.full((300, 400), 128, np.uint8)                      # no candidates at all
    (k0, _), (k1, _) = _surf_both(gpu, flat, "sort", (0, 1), extended=True, keypoints_ratio=0.01)
    assert len(k0) == len(k1) == 0


def test_describe_large_windows_first_identical(gpu, synth_pair_rois):
    roiA, _, _ = synth_pair_rois
    for mode in (0, 1):
        gpu.set_option("describe", mode)
        outs = _surf_both(gpu, roiA, "lpt", (0, 1, 2, 3), extended=True, keypoints_ratio=0.0, hessian_threshold=30.0)
        (k0, d0), (k1, d1) = outs[0], outs[1]
        for k2, d2 in outs[2:]:
            assert np.array_equal(k0, k2) and np.array_equal(d0, d2)
        assert len(k0) > 1000 and (np.floor(21 * k0[:, 2] * np.float32(1.2) / 9) >= 128).sum() > 5
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1)


# ---- tolerance modes (describe = 2, 3): NOT bit-exact; the stated tolerances of DESIGN.md "tolerance modes" are asserted here.
# mode -> (largest |component difference| of any unit-norm descriptor, mean |component difference|, share of the exact path's
#          matches (same query, same train) the mode reproduces)
TOLERANCES = {2: (0.06, 5e-5, 0.995), 3: (0.15, 2e-3, 0.97)}


def _match_set(m):
    return set(map(tuple, np.asarray(m).reshape(-1, 2).tolist()))


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("extended", [True, False])
def test_describe_tolerance_modes_within_stated_tolerance(gpu, synth_pair_rois, extended, mode):
    """describe = 2 blends the gathered 2x2 footprint with fp32 lerps at float positions, describe = 3 takes the window pixels
    from the texture unit's bilinear filter.  Keypoints (position, size, response, orientation) stay identical -- detection
    and orientation do not use the sampler; descriptors, match lists and offsets stay within the stated tolerances."""
    roiA, roiB, true_off = synth_pair_rois
    tol_max, tol_mean, tol_share = TOLERANCES[mode]
    outs = {}
    for md in (1, mode):
        gpu.set_option("describe", md)
        kA, dA = gpu.surf_detect_and_describe(roiA, extended=extended, keypoints_ratio=0.01)
        kB, dB = gpu.surf_detect_and_describe(roiB, extended=extended, keypoints_ratio=0.01)
        m = gpu.match_descriptors(dA, dB, 2, 0.75)
        st, off, votes = gpu.offset_by_mode(kA, kB, m, 3)
        outs[md] = (kA, dA, kB, dB, m, st, off, votes)
    e, t = outs[1], outs[mode]
    assert np.array_equal(e[0], t[0]) and np.array_equal(e[2], t[2])                 # keypoints identical
    for de, dt in ((e[1], t[1]), (e[3], t[3])):
        diff = np.abs(de - dt)
        assert diff.max() <= tol_max, diff.max()
        assert diff.mean() <= tol_mean, diff.mean()
        assert np.allclose(np.linalg.norm(dt, axis=1), 1.0, atol=1e-5)
    me, mt = _match_set(e[4]), _match_set(t[4])
    assert len(me & mt) >= tol_share * len(me), (len(me), len(mt), len(me & mt))
    assert e[5] and t[5] and e[6] == t[6]                                            # same status, identical offset
    assert abs(t[6][0] - true_off[0]) <= 1 and abs(t[6][1] - true_off[1]) <= 1
    print("describe=%d: max |d desc| %.2e mean %.2e identical descriptors %.1f %%, matches %d / %d common %d, votes %d / %d" % (
        mode, max(np.abs(e[1] - t[1]).max(), np.abs(e[3] - t[3]).max()), np.abs(e[1] - t[1]).mean(),
        100.0 * (np.abs(e[1] - t[1]).max(axis=1) == 0).mean(), len(me), len(mt), len(me & mt), e[7], t[7]))


def test_describe_cooperative_giant_windows(gpu):
    """Small batches describe windows of >= 320 px with the warps of a CTA sharing the 21 output rows (surf_describe.cuh,
    cooperative pass) -- same descriptors as the reference sampler and the oracle.  The 1024 x 409 strip holds keypoints of the
    top octave (sizes up to ~200 -> windows up to ~560 px) crossing every border."""
    from imagestitch_b200 import synth
    from oracle import surf
    A, _, _ = synth.pair(seed=11, size=1024, overlap=110, direction=1)
    img = np.ascontiguousarray(A[:409])
    gpu.set_option("describe", 1)
    k1, d1 = gpu.surf_detect_and_describe(img, extended=True, keypoints_ratio=0.0, hessian_threshold=100.0)
    win = (21 * (k1[:, 2] * 1.2 / 9.0)).astype(int)
    assert (win >= 320).sum() >= 3, int((win >= 320).sum())          # the pass has work
    gpu.set_option("describe", 0)
    k0, d0 = gpu.surf_detect_and_describe(img, extended=True, keypoints_ratio=0.0, hessian_threshold=100.0)
    assert np.array_equal(k1, k0) and np.array_equal(d1, d0)
    ko, do = surf.detect_and_compute(img, 100.0, 4, 3, True, False, 0)
    big = win >= 320
    assert np.array_equal(d1[big], do[big]) and np.array_equal(k1[big][:, :4], ko[big][:, :4])
