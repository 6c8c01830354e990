"""JPEG tile decode on the device: bit-identical to cv2.imdecode(data, IMREAD_GRAYSCALE) (Stitcher.py:68-69)."""
import hashlib
import json
import os

import cv2
import numpy as np
import pytest

from test_jpeg_cpu import GOLDEN, SAMPLINGS, _encode, _image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0
    return g


def _cv(data):
    return cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)


@pytest.mark.parametrize("sampling", SAMPLINGS)
@pytest.mark.parametrize("rows,cols,quality,restart", [(75, 131, 35, 0), (409, 517, 92, 7), (64, 64, 100, 1), (1, 1, 80, 0), (7, 9, 60, 0)])
def test_decode_equals_cv2(gpu, sampling, rows, cols, quality, restart):
    data = _encode(_image(rows, cols, rows + quality), quality, sampling, restart)
    assert np.array_equal(gpu.jpeg_decode_gray(data), _cv(data))


def test_single_component_files(gpu):
    for rows, cols in [(16, 16), (17, 33), (300, 1000)]:
        data = _encode(_image(rows, cols, rows, channels=1), 85)
        assert np.array_equal(gpu.jpeg_decode_gray(data), _cv(data))


def test_full_size_tile_batch(gpu):
    """BASELINE config sizes: a batch of 2048 x 2048 tiles, decoded in one call (host entropy threads + one kernel per tile)."""
    datas = [_encode(_image(2048, 2048, 100 + k), 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422) for k in range(5)]
    out = gpu.jpeg_decode_gray(datas)
    assert out.shape == (5, 2048, 2048)
    for k, d in enumerate(datas):
        assert np.array_equal(out[k], _cv(d))


def test_decode_into_device_tile_stack(gpu):
    import torch
    datas = [_encode(_image(200, 300, 7 + k), 90) for k in range(3)]
    stack = torch.full((3, 208, 320), 77, dtype=torch.uint8, device="cuda")
    view = stack[:, 4:204, 10:310]                                   # strided rows and images: decoded in place
    gpu.jpeg_decode_gray_dev(datas, view)
    torch.cuda.synchronize()
    host = stack.cpu().numpy()
    for k, d in enumerate(datas):
        assert np.array_equal(host[k, 4:204, 10:310], _cv(d))
    mask = np.ones_like(host, bool); mask[:, 4:204, 10:310] = False
    assert (host[mask] == 77).all()                                  # nothing written outside the tiles


def test_golden_reference_tiles(gpu):
    cases = json.load(open(os.path.join(GOLDEN, "jpeg_cases.json")))
    for name, c in cases.items():
        if name.startswith("_"):
            continue
        data = np.fromfile(os.path.join(GOLDEN, name), np.uint8).tobytes()
        img = gpu.jpeg_decode_gray(data)
        assert img.shape == (c["rows"], c["cols"])
        assert hashlib.sha256(img.tobytes()).hexdigest() == c["sha256_of_cv2_imdecode_gray"]
        assert np.array_equal(img, _cv(data))


@pytest.mark.parametrize("sampling", SAMPLINGS)
@pytest.mark.parametrize("rows,cols,quality,restart", [(75, 131, 40, 0), (409, 517, 92, 7), (2, 3, 90, 0), (3, 4, 90, 0), (1, 1, 80, 0), (64, 64, 100, 1)])
def test_colour_decode_equals_cv2(gpu, sampling, rows, cols, quality, restart):
    data = _encode(_image(rows, cols, rows + quality), quality, sampling, restart)
    assert np.array_equal(gpu.jpeg_decode_bgr(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))


def test_colour_batch_full_size_and_gray_file(gpu):
    datas = [_encode(_image(1936, 2584, 50 + k), 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422) for k in range(3)]     # the iron / dendritic geometry
    out = gpu.jpeg_decode_bgr(datas)
    for k, d in enumerate(datas):
        assert np.array_equal(out[k], cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR))
    g = _encode(_image(100, 60, 3, channels=1), 80)
    assert np.array_equal(gpu.jpeg_decode_bgr(g), cv2.imdecode(np.frombuffer(g, np.uint8), cv2.IMREAD_COLOR))


def test_golden_reference_tiles_colour(gpu):
    cases = json.load(open(os.path.join(GOLDEN, "jpeg_cases.json")))
    for name, c in cases.items():
        if name.startswith("_"):
            continue
        data = np.fromfile(os.path.join(GOLDEN, name), np.uint8).tobytes()
        img = gpu.jpeg_decode_bgr(data)
        assert img.shape == (c["rows"], c["cols"], 3)
        assert hashlib.sha256(img.tobytes()).hexdigest() == c["sha256_of_cv2_imdecode_color"]


def test_geometry_mismatch_and_unsupported(gpu):
    a = _encode(_image(64, 64, 1), 90); b = _encode(_image(64, 72, 2), 90)
    with pytest.raises(Exception):
        gpu.jpeg_decode_gray([a, b])
    ok, prog = cv2.imencode(".jpg", _image(64, 64, 3), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(gpu.JpegUnsupported):
        gpu.jpeg_decode_gray(prog.tobytes())
