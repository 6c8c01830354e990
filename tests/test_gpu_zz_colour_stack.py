"""GPU: colour twin of the device-resident tile stack (vfsms_tiles_decode_jpeg_bgr / _upload_bgr / vfsms_tiles_mosaic_bgr).  Written
after the round's GPU budget was spent (verified on the CPU emulation): the file name keeps it at the end of the -x run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiles():
    from imagestitch_b200 import synth
    t, offs = synth.tile_sequence(seed=321, n_rows=2, n_cols=3, size=640, overlap=150, noise=1.5)
    return np.stack(t), np.asarray(offs)


def test_colour_twin_one_decode_serves_gray_and_colour(tiles):
    """vfsms_tiles_decode_jpeg_bgr: one entropy-decoding pass fills the gray stack (= cv2.imdecode(data, 0)) and the colour twin
    (= cv2.imdecode(data, IMREAD_COLOR)); the colour mosaic from HBM equals the host-tile mosaic; slots without colour are refused."""
    import cv2
    from imagestitch_b200 import gpu, _lib
    from imagestitch_b200 import sharding as sh
    T, offs = tiles
    n = 4
    rows, cols = T[0].shape
    files = []
    for k, t in enumerate(T[:n]):
        bgr = np.stack([t, np.roll(t, 3, axis=1), 255 - t], axis=-1)
        sampling = (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422)[k % 3]
        files.append(cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, 91, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling])[1])
    files[3] = cv2.imencode(".jpg", T[3], [cv2.IMWRITE_JPEG_QUALITY, 91])[1]              # a single-component file among them
    gpu.tiles_reserve(n + 1, rows, cols)
    gpu.tiles_decode_jpeg_bgr(0, files[:3])
    gpu.tiles_decode_jpeg_bgr(3, files[3:])
    gray = gpu.tiles_download(0, n, rows, cols)
    ref_bgr = []
    for k in range(n):
        assert np.array_equal(gray[k], cv2.imdecode(files[k], cv2.IMREAD_GRAYSCALE)), k
        ref_bgr.append(cv2.imdecode(files[k], cv2.IMREAD_COLOR))
    full = [[0, 0]] + [list(o) for o in offs[:n - 1]]
    origins, rois, shape = sh.rectify_offsets(full, [(rows, cols)] * n)
    for method in ("fadeInAndFadeOut", "notFuse"):
        got = gpu.tiles_mosaic_bgr(0, n, np.asarray(origins, np.int32), rois, np.asarray(full, np.int32), method, shape)
        exp = gpu.mosaic(np.stack(ref_bgr), np.asarray(origins, np.int32), rois, np.asarray(full, np.int32), method, shape)
        assert got.shape == exp.shape == (shape[0], shape[1], 3) and np.array_equal(got, exp), method
    with pytest.raises(_lib.VfsmsError):                                                  # slot 4 holds no colour tile
        gpu.tiles_mosaic_bgr(1, n, np.asarray(origins, np.int32), rois, np.asarray(full, np.int32), "notFuse", shape)
    gpu.tiles_upload_bgr(4, ref_bgr[0])
    gpu.tiles_mosaic_bgr(1, n, np.asarray(origins, np.int32), rois, np.asarray(full, np.int32), "notFuse", shape)
