"""GPU: the device JPEG encoder (vfsms_jpeg_encode_*, jpeg_enc.cu) against cv2.imencode -- the bytes the reference's
cv2.imwrite(..., stitchResult) writes (Stitcher.py:130-131, :196-197) -- and the CPU oracle.  Byte-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return g


def test_bytes_equal_cv2_and_oracle(gpu):
    import cv2
    from oracle import jpeg_encode_oracle as jo
    from test_jpeg_encode_cpu import _images
    for name, im in _images():
        for q in (95, 100, 50, 10):
            ref = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
            out = gpu.jpeg_encode(im, q)
            assert out == ref, (name, q, len(out), len(ref))
        assert gpu.jpeg_encode(im) == jo.encode(im), name


def test_strided_views_and_capacity_retry(gpu):
    import cv2
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=9, size=384, overlap=40, direction=1)
    col = np.dstack([A, np.roll(A, 3, 0), np.roll(A, 4, 1)])
    for view in (A[5:205, 17:300], col[5:205, 17:300], A[::2, ::2], col[:, ::-1]):
        assert gpu.jpeg_encode(view) == cv2.imencode(".jpg", np.ascontiguousarray(view))[1].tobytes()
    noise = np.random.default_rng(0).integers(0, 256, (256, 256, 3), dtype=np.uint8)     # q100 noise: larger than the first buffer guess
    assert gpu.jpeg_encode(noise, 100) == cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_QUALITY, 100])[1].tobytes()


def test_argument_errors(gpu):
    from imagestitch_b200 import _lib
    with pytest.raises(TypeError):
        gpu.jpeg_encode(np.zeros((8, 8), np.float32))
    with pytest.raises(TypeError):
        gpu.jpeg_encode(np.zeros((8, 8, 4), np.uint8))
    with pytest.raises(_lib.VfsmsError):
        gpu.jpeg_encode(np.zeros((1, 70000), np.uint8))                                 # JPEG dimensions are 16-bit


def test_mosaic_sized_canvas_round_trip(gpu):
    """a canvas of mosaic proportions (odd sizes, empty = 0 regions): identical to cv2 at full size, and the library's own decoder
    reads the file back to what cv2 reads (encode -> decode closes over both directions of the data format)"""
    import cv2
    from imagestitch_b200 import synth
    A, B, _ = synth.pair(seed=21, size=1024, overlap=100, direction=1)
    canvas = np.zeros((1531, 2071), np.uint8)
    canvas[:1024, :1024] = A; canvas[500:1524, 1040:2064] = B
    col = np.dstack([canvas, np.roll(canvas, 11, 1), canvas[::-1]])
    for img in (canvas, col):
        data = gpu.jpeg_encode(img)
        assert data == cv2.imencode(".jpg", img)[1].tobytes()
        flag = cv2.IMREAD_GRAYSCALE if img.ndim == 2 else cv2.IMREAD_COLOR
        back = gpu.jpeg_decode_gray(data) if img.ndim == 2 else gpu.jpeg_decode_bgr(data)
        assert np.array_equal(back, cv2.imdecode(np.frombuffer(data, np.uint8), flag))
        # lossy, but close: q95 gray; the colour planes here are unrelated images, 4:2:0 costs more
        assert np.abs(back.astype(int) - img.astype(int)).mean() < (2.0 if img.ndim == 2 else 25.0)


def test_stitcher_writes_identical_files(gpu, tmp_path):
    """imageSetStitchWithMutiple(..., outputfileExtension="jpg") (Main.py:20): the result file equals cv2.imwrite's, gray and colour"""
    import cv2
    from imagestitch_b200 import Stitcher as S
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=4, size=256, overlap=30, direction=1)
    col = np.dstack([A, np.roll(A, 2, 0), np.roll(A, 2, 1)])
    for k, img in enumerate((A, col)):
        p1, p2 = str(tmp_path / ("a%d.jpg" % k)), str(tmp_path / ("b%d.jpg" % k))
        assert S._imwrite(p1, img, "b200") is True and cv2.imwrite(p2, img)
        assert open(p1, "rb").read() == open(p2, "rb").read()
    p3 = str(tmp_path / "c.png")
    assert S._imwrite(p3, A, "b200") and np.array_equal(cv2.imread(p3, 0), A)           # other formats: cv2
    assert S._imwrite(os.path.join(str(tmp_path), "missing_dir", "x.jpg"), A, "b200") is False
