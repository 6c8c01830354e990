"""Device-resident tile stack: decode once, align ROI strips in place, mosaic from the stack -- results identical to the
host-array entry points (and therefore to the reference path they are tested against)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiles():
    from imagestitch_b200 import synth
    t, offs = synth.tile_sequence(seed=321, n_rows=2, n_cols=3, size=640, overlap=150, noise=1.5)
    return np.stack(t), np.asarray(offs)


@pytest.mark.parametrize("direction", [1, 2, 3, 4])
@pytest.mark.parametrize("i", [1, 2])
def test_tiles_align_equals_host_rois(tiles, direction, i):
    from imagestitch_b200 import gpu
    from imagestitch_b200.ImageUtility import Method
    T, _ = tiles
    m = Method()
    n, H, W = T.shape
    gpu.tiles_reserve(n, H, W)
    gpu.tiles_upload(0, T)
    assert np.array_equal(gpu.tiles_download(0, n, H, W), T)
    ratio = 0.2 * i
    L = int(np.floor((H if direction in (1, 3) else W) * ratio))
    params = gpu.surf_params()
    got = gpu.tiles_align(0, n - 1, direction, L, params=params)
    A = np.stack([np.ascontiguousarray(m.getROIRegionForIncreMethod(T[k], direction, "first", ratio)) for k in range(n - 1)])
    B = np.stack([np.ascontiguousarray(m.getROIRegionForIncreMethod(T[k + 1], direction, "second", ratio)) for k in range(n - 1)])
    ref = gpu.align_batch(A, B, params=params)
    for name in ("status", "d_row", "d_col", "votes", "n_a", "n_b", "n_matches"):
        if name in ref.dtype.names:
            assert np.array_equal(got[name], ref[name]), name
    # a sub-range of the stack
    sub = gpu.tiles_align(2, 2, direction, L, params=params)
    assert np.array_equal(sub["d_row"], ref["d_row"][2:4]) and np.array_equal(sub["status"], ref["status"][2:4])


def test_tiles_mosaic_equals_host_mosaic(tiles):
    from imagestitch_b200 import gpu
    T, _ = tiles
    n, H, W = 3, T.shape[1], T.shape[2]
    gpu.tiles_reserve(n, H, W)
    gpu.tiles_upload(0, T[:n])
    origins = np.array([[0, 0], [5, W - 150], [2, 2 * (W - 150) + 7]], np.int32)
    pair = np.array([[0, 0], [5, W - 150], [-3, W - 143]], np.int32)
    rois = np.zeros((n, 4), np.int32)
    for k in range(1, n):
        rois[k] = (max(origins[k, 0], 0), origins[k, 1], min(origins[k, 0] + H, H + 5), origins[k - 1, 1] + W)
    shape = (H + 8, int(origins[-1, 1]) + W)
    for method in ("notFuse", "average", "fadeInAndFadeOut", "trigonometric"):
        a = gpu.tiles_mosaic(0, n, origins, rois, pair, method, shape)
        b = gpu.mosaic(T[:n], origins, rois, pair, method, shape)
        assert np.array_equal(a, b), method


def test_stitcher_on_jpeg_tiles_device_stack_equals_cv2_decode(tiles, tmp_path):
    """What Main.py does on a directory of JPEG tiles: library decode + tile stack vs cv2 decode + host ROIs -- same
    offsets, same mosaic bytes."""
    import cv2
    from Stitcher import Stitcher
    T, offs = tiles
    d = tmp_path / "set" / "1"
    d.mkdir(parents=True)
    for k, t in enumerate(T):
        cv2.imwrite(str(d / ("tile-%02d.jpg" % k)), t, [cv2.IMWRITE_JPEG_QUALITY, 95])
    results = {}
    for decoder in ("b200", "cv2"):
        Stitcher.featureMethod = "surf"; Stitcher.isColorMode = False; Stitcher.isGPUAvailable = False; Stitcher.isEnhance = False
        Stitcher.searchRatio = 0.75; Stitcher.offsetCaculate = "mode"; Stitcher.offsetEvaluate = 3; Stitcher.roiRatio = 0.2
        Stitcher.fuseMethod = "fadeInAndFadeOut"; Stitcher.direction = 1; Stitcher.directIncre = 1; Stitcher.isPrintLog = False
        Stitcher.decoder = decoder
        st = Stitcher()
        seen = {}
        orig = st.getStitchByOffset

        def spy(fileList, offsetList, _orig=orig, _seen=seen):
            _seen["offsets"] = [list(o) for o in offsetList]
            return _orig(fileList, offsetList)
        st.getStitchByOffset = spy
        out = tmp_path / ("out_" + decoder)
        st.imageSetStitchWithMutiple(str(tmp_path / "set"), str(out), 1, st.calculateOffsetForFeatureSearchIncre, fileExtension="jpg",
                                     outputfileExtension="png")
        results[decoder] = (seen["offsets"], cv2.imread(os.path.join(str(out), "stitching_result_1.png"), 0))
    Stitcher.decoder = "b200"; Stitcher.isPrintLog = True; Stitcher.isColorMode = True; Stitcher.fuseMethod = "notFuse"; Stitcher.direction = 1
    assert results["b200"][0] == results["cv2"][0]
    assert len(results["b200"][0]) == len(offs)
    for got, true in zip(results["b200"][0], offs):
        assert abs(got[0] - true[0]) <= 1 and abs(got[1] - true[1]) <= 1
    assert results["b200"][1] is not None and np.array_equal(results["b200"][1], results["cv2"][1])


def test_stitcher_colour_mode_on_jpeg_tiles(tiles, tmp_path):
    """Main.py's default (isColorMode = True): gray decode of the colour JPEGs for alignment, colour decode for the mosaic --
    both through the library, same result as with cv2 decoding."""
    import cv2
    from Stitcher import Stitcher
    T, offs = tiles
    d = tmp_path / "set" / "1"
    d.mkdir(parents=True)
    for k, t in enumerate(T[:4]):
        bgr = np.stack([t, np.roll(t, 3, axis=1), 255 - t], axis=-1)
        cv2.imwrite(str(d / ("tile-%02d.jpg" % k)), bgr, [cv2.IMWRITE_JPEG_QUALITY, 93])
    results = {}
    for decoder in ("b200", "cv2"):
        Stitcher.featureMethod = "surf"; Stitcher.isColorMode = True; Stitcher.isGPUAvailable = False; Stitcher.isEnhance = False
        Stitcher.searchRatio = 0.75; Stitcher.offsetCaculate = "mode"; Stitcher.offsetEvaluate = 3; Stitcher.roiRatio = 0.2
        Stitcher.fuseMethod = "fadeInAndFadeOut"; Stitcher.direction = 1; Stitcher.directIncre = 1; Stitcher.isPrintLog = False
        Stitcher.decoder = decoder
        st = Stitcher()
        out = tmp_path / ("out_" + decoder)
        st.imageSetStitchWithMutiple(str(tmp_path / "set"), str(out), 1, st.calculateOffsetForFeatureSearchIncre, fileExtension="jpg",
                                     outputfileExtension="png")
        names = sorted(os.listdir(str(out)))
        results[decoder] = [cv2.imread(os.path.join(str(out), n), cv2.IMREAD_COLOR) for n in names]
    Stitcher.decoder = "b200"; Stitcher.isPrintLog = True; Stitcher.fuseMethod = "notFuse"; Stitcher.direction = 1
    assert len(results["b200"]) == len(results["cv2"]) >= 1
    for a, b in zip(results["b200"], results["cv2"]):
        assert a is not None and a.ndim == 3 and np.array_equal(a, b)


def test_tiles_align_list_equals_per_direction_calls():
    from imagestitch_b200 import gpu
    """vfsms_tiles_align_list: arbitrary (pair, direction) lists in one fused call = the per-direction contiguous calls, and the
    strided variant = every step-th pair of them (the calls a sharded run's batched search makes, sharding.py)."""
    from imagestitch_b200 import synth
    tiles, _ = synth.tile_sequence(31, 2, 3, size=512, overlap=64)             # 6 tiles, serpentine: directions 2, 2, 1, 4, 4
    gpu.tiles_reserve(len(tiles), 512, 512)
    gpu.tiles_upload(0, tiles)
    L = int(0.2 * 512)
    ref = {d: gpu.tiles_align(0, 5, d, L) for d in (1, 2, 3, 4)}
    assert ref[2]["status"][[0, 1]].all() and ref[1]["status"][2] and ref[4]["status"][[3, 4]].all()
    for dirs in ((1, 3), (2, 4)):
        pairs = [4, 0, 2, 2, 3, 1, 0]
        ds = [dirs[k % 2] for k in range(len(pairs))]
        got = gpu.tiles_align_list(pairs, ds, L)
        for k, (p, d) in enumerate(zip(pairs, ds)):
            assert got[k] == ref[d][p], (k, p, d)
    strided = gpu.tiles_align(0, 3, 2, L, step=2)
    assert all(strided[k] == ref[2][2 * k] for k in range(3))
    with pytest.raises(gpu.VfsmsError):
        gpu.tiles_align_list([0, 1], [1, 2], L)                                  # mixed strip shapes


def _small_stack():
    from imagestitch_b200 import gpu, synth
    tiles, _ = synth.tile_sequence(31, 2, 3, size=512, overlap=64)
    tiles = np.stack(tiles)
    gpu.tiles_reserve(len(tiles), 512, 512)
    gpu.tiles_upload(0, tiles)


_CALLS = [([0, 1, 3], 1, [1, 3, 1]), ([0, 1, 3, 4], 1, [2, 4, 2, 4]), ([2], 2, [1]), ([2, 3], 2, [4, 2])]


def test_second_context_borrows_the_stack():
    """vfsms_tiles_attach: a second context of the device reads lane 0's stack and gives lane 0's results."""
    from imagestitch_b200 import gpu, sharding
    _small_stack()
    ev = sharding.tiles_batch_evaluator(0, lambda i, d: int(i * 0.2 * 512), lanes=1)
    gpu.tiles_attach(1)
    for c in _CALLS:
        assert np.array_equal(ev(*c, lane=1), ev(*c))
    _small_stack()                      # lane 0 re-reserves: attach again, same answers
    gpu.tiles_attach(1)
    assert np.array_equal(ev(*_CALLS[1], lane=1), ev(*_CALLS[1]))


def test_two_contexts_run_a_round_side_by_side():
    """The two strip shapes of a search round run from two host threads on two contexts
    (sharding.tiles_batch_evaluator(...).many) and give the results of running them one after the other."""
    from imagestitch_b200 import sharding
    _small_stack()
    ev = sharding.tiles_batch_evaluator(0, lambda i, d: int(i * 0.2 * 512), lanes=2)
    for _ in range(3):
        for c, got in zip(_CALLS, ev.many(_CALLS)):
            assert np.array_equal(got, ev(*c))
    # a sharded search through _run_requests uses .many for rounds with two shapes: same table as a one-lane evaluator
    ev1 = sharding.tiles_batch_evaluator(0, lambda i, d: int(i * 0.2 * 512), lanes=1)
    t2, _ = sharding.evaluate_shard_batched(ev, 0, 5, 1, 1, 0.2)
    t1, _ = sharding.evaluate_shard_batched(ev1, 0, 5, 1, 1, 0.2)
    assert np.array_equal(t1, t2)
