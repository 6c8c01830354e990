"""GPU: Huffman decoding of JPEG tiles on the device (vfsms_set_option "entropy" = 1: self-synchronising parallel decode of
1024-bit subsequences, jpeg.cu) -- same pixels as the host stage and as cv2.imdecode, file by file.  Written after the round's
GPU budget was spent (verified on the CPU emulation): the file name keeps it at the end of the -x run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    default = g.get_option("entropy")
    g.set_option("entropy", 1)
    yield g
    g.set_option("entropy", default)


def _content():
    from imagestitch_b200 import synth
    A, _, _ = synth.pair(seed=3, size=512, overlap=60, direction=1)
    return A, np.dstack([A, np.roll(A, 7, 0), 255 - np.roll(A, 5, 1)])


def test_every_sampling_quality_and_size_equals_cv2(gpu):
    import cv2
    A, col = _content()
    S = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
         cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]
    passes = []
    for (h, w) in ((1, 1), (8, 8), (17, 33), (131, 75), (409, 517)):
        for q in (35, 92, 100):
            for samp in S:
                for img in (A[:h, :w], col[:h, :w]):
                    data = cv2.imencode(".jpg", np.ascontiguousarray(img), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp])[1]
                    assert np.array_equal(gpu.jpeg_decode_gray(data), cv2.imdecode(data, cv2.IMREAD_GRAYSCALE)), (h, w, q, samp, img.ndim)
                    assert np.array_equal(gpu.jpeg_decode_bgr(data), cv2.imdecode(data, cv2.IMREAD_COLOR)), (h, w, q, samp, img.ndim)
                    passes.append(gpu.jpeg_last_entropy_passes())
    assert min(passes) >= 4 and max(passes) < 400          # the decoders re-synchronise: far fewer passes than subsequences


def test_batches_tile_stack_and_host_stage_identity(gpu):
    import cv2
    A, col = _content()
    files = [cv2.imencode(".jpg", np.roll(col, 17 * k, 0), [cv2.IMWRITE_JPEG_QUALITY, 60 + 5 * k])[1] for k in range(7)]
    files[3] = cv2.imencode(".jpg", np.roll(A, 5, 1), [cv2.IMWRITE_JPEG_QUALITY, 88])[1]        # a single-component file in the batch
    dev = gpu.jpeg_decode_bgr(files)
    gpu.set_option("entropy", 0)
    host = gpu.jpeg_decode_bgr(files)
    gpu.set_option("entropy", 1)
    assert np.array_equal(dev, host)
    for k in range(7):
        assert np.array_equal(dev[k], cv2.imdecode(files[k], cv2.IMREAD_COLOR)), k
    # the tile stack: gray slots and the colour twin from one pass
    n, (rows, cols) = len(files), A.shape
    gpu.tiles_reserve(n, rows, cols)
    gpu.tiles_decode_jpeg_bgr(0, files)
    gray = gpu.tiles_download(0, n, rows, cols)
    for k in range(n):
        assert np.array_equal(gray[k], cv2.imdecode(files[k], cv2.IMREAD_GRAYSCALE)), k
    gpu.tiles_decode_jpeg(0, files[:3])
    assert np.array_equal(gpu.tiles_download(0, 3, rows, cols), gray[:3])


def test_restart_intervals_and_damaged_files_take_the_host_stage(gpu):
    """files the device stage does not decode itself (DRI; streams that end before the last block) give the host stage's result"""
    import cv2
    A, col = _content()
    rst = cv2.imencode(".jpg", col[:200, :264], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 5])[1]
    assert np.array_equal(gpu.jpeg_decode_bgr(rst), cv2.imdecode(rst, cv2.IMREAD_COLOR))
    good = cv2.imencode(".jpg", col[:200, :264], [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
    cut = good[:len(good) * 2 // 3].copy()                                                # truncated in the middle of the scan
    noisy = good.copy()
    rng = np.random.default_rng(5)
    idx = rng.integers(700, len(noisy) - 2, 12)
    noisy[idx] = rng.integers(0, 255, 12).astype(np.uint8)                                # damaged entropy-coded data
    for data in (cut, noisy):
        try:
            dev = gpu.jpeg_decode_bgr(data)
        except gpu.VfsmsError:
            dev = None
        gpu.set_option("entropy", 0)
        try:
            host = gpu.jpeg_decode_bgr(data)
        except gpu.VfsmsError:
            host = None
        gpu.set_option("entropy", 1)
        assert (dev is None) == (host is None)
        if dev is not None:
            assert np.array_equal(dev, host)


def test_large_tile_batch(gpu):
    """2048^2 tiles as bench.py ingests them (q92): identical to cv2, a handful of synchronisation passes"""
    import cv2
    from imagestitch_b200 import synth
    T, _ = synth.tile_sequence(seed=12, n_rows=1, n_cols=3, size=2048, overlap=205)
    files = [cv2.imencode(".jpg", t, [cv2.IMWRITE_JPEG_QUALITY, 92])[1] for t in T]
    out = gpu.jpeg_decode_gray(files)
    for k, f in enumerate(files):
        assert np.array_equal(out[k], cv2.imdecode(f, cv2.IMREAD_GRAYSCALE)), k
    assert gpu.jpeg_last_entropy_passes() <= 64


def test_golden_reference_tiles(gpu, golden_dir):
    """two of the reference's demo micrographs (camera / microscope software encoders, not cv2's): SHA-256 of cv2's decode"""
    import hashlib
    import json
    import os
    cases = json.load(open(os.path.join(golden_dir, "jpeg_cases.json")))
    for name, c in cases.items():
        if name.startswith("_"):
            continue
        data = np.fromfile(os.path.join(golden_dir, name), np.uint8).tobytes()
        assert hashlib.sha256(gpu.jpeg_decode_gray(data).tobytes()).hexdigest() == c["sha256_of_cv2_imdecode_gray"], name
        assert hashlib.sha256(gpu.jpeg_decode_bgr(data).tobytes()).hexdigest() == c["sha256_of_cv2_imdecode_color"], name
