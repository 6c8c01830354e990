"""GPU: the whole reference workflow as Main.py runs it -- JPEG tiles in a directory -> offsets -> colour fade mosaic -> .jpg result --
through Stitcher.imageSetStitchWithMutiple.  The result FILE must be byte-identical whether tiles are decoded and the result encoded by
the library (host or device entropy stage) or by cv2: decode (gray + colour twin), alignment on the HBM stack, colour mosaic from HBM
and the encoder all sit on that path.  Written after the round's GPU budget was spent (verified on the CPU emulation)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(root, out, decoder, encoder, entropy):
    from imagestitch_b200 import gpu
    from Stitcher import Stitcher
    Stitcher.featureMethod = "surf"; Stitcher.isColorMode = True; Stitcher.isGPUAvailable = False; Stitcher.isEnhance = False
    Stitcher.searchRatio = 0.75; Stitcher.offsetCaculate = "mode"; Stitcher.offsetEvaluate = 3; Stitcher.roiRatio = 0.2
    Stitcher.fuseMethod = "fadeInAndFadeOut"; Stitcher.direction = 1; Stitcher.directIncre = 1; Stitcher.isPrintLog = False
    Stitcher.decoder = decoder; Stitcher.encoder = encoder
    default_entropy = gpu.get_option("entropy")
    gpu.set_option("entropy", entropy)
    try:
        st = Stitcher()
        st.imageSetStitchWithMutiple(root, out, 1, st.calculateOffsetForFeatureSearchIncre, fileExtension="jpg", outputfileExtension="jpg")
    finally:
        gpu.set_option("entropy", default_entropy)
        Stitcher.decoder = "b200"; Stitcher.encoder = "b200"; Stitcher.isPrintLog = True; Stitcher.fuseMethod = "notFuse"; Stitcher.direction = 1
    names = sorted(os.listdir(out))
    return {n: open(os.path.join(out, n), "rb").read() for n in names}


def test_result_file_identical_for_library_and_cv2_codecs(tmp_path):
    import cv2
    from imagestitch_b200 import gpu, synth
    assert gpu.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    tiles, offs = synth.tile_sequence(seed=77, n_rows=2, n_cols=2, size=448, overlap=80, noise=1.5)
    d = tmp_path / "set" / "1"
    d.mkdir(parents=True)
    for k, t in enumerate(tiles):
        cv2.imwrite(str(d / ("tile-%02d.jpg" % k)), np.stack([t, np.roll(t, 2, axis=1), 255 - t // 2], axis=-1), [cv2.IMWRITE_JPEG_QUALITY, 92])
    root = str(tmp_path / "set")
    lib = _run(root, str(tmp_path / "out_lib"), "b200", "b200", 0)
    ref = _run(root, str(tmp_path / "out_cv2"), "cv2", "cv2", 0)
    dev = _run(root, str(tmp_path / "out_dev"), "b200", "b200", 1)
    assert list(lib) == list(ref) == list(dev) == ["stitching_result_1.jpg"]
    assert lib["stitching_result_1.jpg"] == ref["stitching_result_1.jpg"] == dev["stitching_result_1.jpg"]
    img = cv2.imdecode(np.frombuffer(lib["stitching_result_1.jpg"], np.uint8), cv2.IMREAD_COLOR)
    assert img is not None and img.shape[0] > 700 and img.shape[1] > 700 and (img.sum(axis=2) > 0).mean() > 0.9      # all four tiles placed


def test_colour_twin_that_does_not_fit_falls_back_to_gray_stack(tmp_path, monkeypatch):
    """The colour twin is 3 x the gray stack; when reserving / filling it fails (VfsmsError) the sequence is loaded into the gray stack alone
    and the mosaic decodes the colour tiles batch by batch -- the result file must not change."""
    import cv2
    from imagestitch_b200 import gpu, synth
    import imagestitch_b200.Stitcher as S
    assert gpu.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    tiles, offs = synth.tile_sequence(seed=78, n_rows=1, n_cols=3, size=384, overlap=72, noise=1.5)
    d = tmp_path / "set" / "1"
    d.mkdir(parents=True)
    for k, t in enumerate(tiles):
        cv2.imwrite(str(d / ("tile-%02d.jpg" % k)), np.stack([t, np.roll(t, 3, axis=0), 255 - t // 3], axis=-1), [cv2.IMWRITE_JPEG_QUALITY, 90])
    root = str(tmp_path / "set")
    full = _run(root, str(tmp_path / "out_full"), "b200", "b200", 0)
    assert S._last_sequence_has_color

    def no_room(first, datas, device=0):
        raise gpu.VfsmsError("vfsms_tiles_decode_jpeg_bgr failed (-2): out of device memory (test)")
    monkeypatch.setattr(gpu, "tiles_decode_jpeg_bgr", no_room)
    gray_only = _run(root, str(tmp_path / "out_gray"), "b200", "b200", 0)
    assert not S._last_sequence_has_color
    assert list(full) == list(gray_only) == ["stitching_result_1.jpg"]
    assert full["stitching_result_1.jpg"] == gray_only["stitching_result_1.jpg"]
