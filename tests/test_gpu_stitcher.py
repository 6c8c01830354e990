"""End-to-end through the reference's call surface (what Main.py does) on a synthetic serpentine tile set."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tile_set(tmp_path_factory):
    import cv2
    from imagestitch_b200 import synth
    tiles, offs = synth.tile_sequence(seed=123, n_rows=2, n_cols=3, size=512, overlap=80, noise=1.5)
    root = tmp_path_factory.mktemp("proj")
    d = root / "1"
    d.mkdir()
    for k, t in enumerate(tiles):
        cv2.imwrite(str(d / ("t-%03d.Png" % k)), t)                      # mixed-case extension like demoImages/zirconCL
    return str(root), tiles, offs


def _configure(Stitcher):
    Stitcher.featureMethod = "surf"; Stitcher.isColorMode = False; Stitcher.isGPUAvailable = False
    Stitcher.searchRatio = 0.75; Stitcher.offsetCaculate = "mode"; Stitcher.offsetEvaluate = 3; Stitcher.roiRatio = 0.2
    Stitcher.fuseMethod = "fadeInAndFadeOut"; Stitcher.direction = 1; Stitcher.directIncre = 1; Stitcher.isPrintLog = False


def test_image_set_stitch_like_main(tile_set, tmp_path):
    import cv2
    from Stitcher import Stitcher
    root, tiles, offs = tile_set
    _configure(Stitcher)
    st = Stitcher()
    seen = {}
    orig = st.getStitchByOffset

    def spy(fileList, offsetList):
        seen["offsets"] = [list(o) for o in offsetList]
        return orig(fileList, offsetList)
    st.getStitchByOffset = spy
    out_dir = str(tmp_path / "result") + "\\"                           # Windows-style separators as in Main.py:19
    st.imageSetStitchWithMutiple(root.replace("/", "\\"), out_dir, 1, st.calculateOffsetForFeatureSearchIncre,
                                 startNum=1, fileExtension="png", outputfileExtension="png")
    assert len(seen["offsets"]) == len(offs)
    for got, true in zip(seen["offsets"], offs):
        assert abs(got[0] - true[0]) <= 1 and abs(got[1] - true[1]) <= 1, (seen["offsets"], offs.tolist())
    res = cv2.imread(os.path.join(str(tmp_path / "result"), "stitching_result_1.png"), 0)
    assert res is not None and res.shape[0] > 900 and res.shape[1] > 1300 and (res > 0).mean() > 0.9
    # the batched flow returns exactly what the pairwise method returns pair by pair (carried direction included)
    Stitcher.direction = 1
    st2 = Stitcher()
    seq = []
    for k in range(len(tiles) - 1):
        s, o = st2.calculateOffsetForFeatureSearchIncre([tiles[k], tiles[k + 1]])
        assert s
        seq.append(o)
    assert seq == seen["offsets"]
    Stitcher.direction = 1; Stitcher.isPrintLog = True; Stitcher.isColorMode = True; Stitcher.fuseMethod = "notFuse"


def test_full_frame_method_and_cache(tile_set):
    from Stitcher import Stitcher
    root, tiles, offs = tile_set
    _configure(Stitcher)
    st = Stitcher()
    st.tempImageFeature.isBreak = True
    s, o = st.calculateOffsetForFeatureSearch([tiles[0], tiles[1]])
    assert s and abs(o[0] - offs[0][0]) <= 1 and abs(o[1] - offs[0][1]) <= 1
    assert st.tempImageFeature.isBreak is False and st.tempImageFeature.feature is not None
    s, o = st.calculateOffsetForFeatureSearch([tiles[1], tiles[2]])        # reuses the cached features of tile 1
    assert s and abs(o[0] - offs[1][0]) <= 1 and abs(o[1] - offs[1][1]) <= 1
    st.tempImageFeature.isBreak = True
    Stitcher.isPrintLog = True


def test_phase_incre_equals_cv2_evaluation(tile_set):
    """Same (status, offset) as the reference function body evaluated with cv2.phaseCorrelate (Stitcher.py:205-258)."""
    import cv2
    from Stitcher import Stitcher
    root, tiles, offs = tile_set
    _configure(Stitcher)
    Stitcher.directIncre = 1; Stitcher.direction = 1
    st = Stitcher()
    got = st.calculateOffsetForPhaseCorrleateIncre([tiles[0], tiles[1]])
    # reference loop with cv2
    A, B = tiles[0], tiles[1]
    exp = (False, None)
    d = 1
    for i in range(1, int(np.floor(0.5 / 0.2) + 1) + 1):
        status = False
        ini = 1 if i == 1 else d
        d = ini
        while True:
            ra = st.getROIRegionForIncreMethod(A, d, "first", i * 0.2); rb = st.getROIRegionForIncreMethod(B, d, "second", i * 0.2)
            (sh, resp) = cv2.phaseCorrelate(np.float64(ra), np.float64(rb))
            off = [int(sh[1]), int(sh[0])]
            if resp > 0.15:
                status = True
                break
            d = st.directionIncrease(d)
            if d == ini:
                break
        if status:
            exp = (True, st._roi_origin_back(off, [A, B], i, d))
            break
    if exp[0]:
        assert got[0] and got[1] == exp[1]
    else:
        assert got[0] is False
    Stitcher.direction = 1; Stitcher.isPrintLog = True
    if "direction" in st.__dict__:
        del st.__dict__["direction"]


def test_enhancement_equals_cv2(tile_set):
    """isEnhance path (Stitcher.py:269-276): equalizeHist and CLAHE(20, 5x5) vs cv2, incl. sizes not divisible by the grid."""
    import cv2
    from imagestitch_b200 import gpu
    root, tiles, offs = tile_set
    # 100 x 103 / 103 x 100 / 100 x 100: one side divisible by the 5 x 5 grid, the other not (cv2 then pads BOTH), and both divisible
    for img in (tiles[0], tiles[1][:103, :257], np.ascontiguousarray(tiles[2][:, 512 - 102:]), tiles[0][:100, :103], tiles[1][:103, :100],
                tiles[2][:100, :100]):
        assert np.array_equal(gpu.enhance(img, clahe=False), cv2.equalizeHist(np.ascontiguousarray(img)))
        ref = cv2.createCLAHE(clipLimit=20, tileGridSize=(5, 5)).apply(np.ascontiguousarray(img))
        out = gpu.enhance(img, clahe=True, clip_limit=20, tile_size=5)
        d = np.abs(out.astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, (int(d.max()), float((d > 0).mean()))
    flat = np.full((40, 50), 9, np.uint8)
    assert np.array_equal(gpu.enhance(flat), cv2.equalizeHist(flat))
