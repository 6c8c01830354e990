"""JPEG tile decode, CPU side: the oracle restatement is pinned against cv2.imdecode (the reference's decoder call,
Stitcher.py:68-69), and the library's host entropy stage is checked against the oracle.  No GPU needed."""
import hashlib
import json
import os

import cv2
import numpy as np
import pytest

from oracle import jpeg_oracle as jo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SAMPLINGS = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
             cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]


def _image(rows, cols, seed, channels=3):
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.standard_normal((rows + 8, cols + 8, channels)).astype(np.float32), (0, 0), 2.0) * 300 + 128
    img = img.reshape(rows + 8, cols + 8, channels)
    img += rng.standard_normal(img.shape).astype(np.float32) * 6
    img = img[:rows, :cols].clip(0, 255).astype(np.uint8)
    return np.ascontiguousarray(img[:, :, 0] if channels == 1 else img)


def _encode(img, quality=90, sampling=None, restart=0):
    p = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_RST_INTERVAL, restart]
    if sampling is not None:
        p += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling]
    ok, buf = cv2.imencode(".jpg", img, p)
    assert ok
    return buf.tobytes()


@pytest.mark.parametrize("sampling", SAMPLINGS)
@pytest.mark.parametrize("quality,restart", [(35, 0), (92, 5), (100, 1)])
def test_oracle_equals_cv2(sampling, quality, restart):
    data = _encode(_image(75, 131, quality + restart), quality, sampling, restart)
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(jo.decode_gray(data), ref)


def test_oracle_grayscale_file_and_tiny_sizes():
    for rows, cols in [(1, 1), (7, 9), (8, 8), (16, 16), (17, 33)]:
        data = _encode(_image(rows, cols, rows * 100 + cols, channels=1), 80)
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(jo.decode_gray(data), ref)


@pytest.mark.parametrize("sampling", SAMPLINGS)
@pytest.mark.parametrize("rows,cols", [(75, 131), (33, 17), (2, 3), (3, 4), (5, 5), (1, 2), (16, 15)])
def test_colour_oracle_equals_cv2(sampling, rows, cols):
    """IMREAD_COLOR path (Stitcher.py:382,401): fancy upsampling incl. the downsampled_width <= 2 rule, YCbCr tables."""
    data = _encode(_image(rows, cols, rows * 7 + cols), 90, sampling)
    assert np.array_equal(jo.decode_bgr(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))


def test_host_stage_all_components_equal_oracle():
    from imagestitch_b200 import gpu
    data = _encode(_image(95, 123, 5), 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, restart=2)
    info, coefs = jo.all_coefficients(data)
    for c in range(3):
        coef, quant, hv = gpu.jpeg_component_coefficients(data, c)
        assert np.array_equal(coef, coefs[c]) and hv == (info["comps"][c][1], info["comps"][c][2])
        assert np.array_equal(quant.astype(np.int32), info["quant"][info["comps"][c][3]])


def test_range_limit_wraps_like_libjpeg():
    x = np.arange(-2048, 2048)
    idx = x & 1023
    table = np.concatenate([np.arange(128, 256), np.full(384, 255), np.zeros(384, np.int64), np.arange(0, 128)])   # jdmaster.c prepare_range_limit_table
    assert np.array_equal(jo.range_limit(x), table[idx].astype(np.uint8))


@pytest.mark.parametrize("sampling", SAMPLINGS)
def test_host_entropy_stage_equals_oracle(sampling):
    from imagestitch_b200 import gpu
    data = _encode(_image(203, 157, 11), 88, sampling, restart=3)
    info, coef_o = jo.luma_coefficients(data)
    coef, quant = gpu.jpeg_luma_coefficients(data)
    assert coef.shape == coef_o.shape and np.array_equal(coef, coef_o)
    assert np.array_equal(quant.astype(np.int32), info["quant"][info["comps"][0][3]])
    assert gpu.jpeg_info(data) == (203, 157, 3)


def test_golden_files_through_host_stage_and_oracle_idct():
    """The reference's own tiles (custom Huffman / quantisation tables of the microscope software)."""
    from imagestitch_b200 import gpu
    cases = json.load(open(os.path.join(GOLDEN, "jpeg_cases.json")))
    assert cases["_demo_sweep"]["files"] == cases["_demo_sweep"]["bit_exact_vs_cv2"] == cases["_demo_sweep"]["color_bit_exact_vs_cv2"] == 140
    for name, c in cases.items():
        if name.startswith("_"):
            continue
        data = np.fromfile(os.path.join(GOLDEN, name), np.uint8).tobytes()
        coef, quant = gpu.jpeg_luma_coefficients(data)
        img = jo.idct_islow(coef, quant.astype(np.int32))[:c["rows"], :c["cols"]]
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == c["sha256_of_cv2_imdecode_gray"]
        assert np.array_equal(img, cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE))


def test_unsupported_and_damaged_input():
    from imagestitch_b200 import gpu
    img = _image(64, 64, 3)
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(gpu.JpegUnsupported):
        gpu.jpeg_info(prog.tobytes())
    with pytest.raises(gpu.JpegUnsupported):
        gpu.jpeg_info(b"\x89PNG\r\n\x1a\n" + bytes(64))
    data = _encode(img, 90)
    # a truncated entropy segment decodes to *something* of the right geometry without reading out of bounds (libjpeg warns and pads)
    coef, _ = gpu.jpeg_luma_coefficients(data[: len(data) // 2])
    assert coef.shape == (8, 8, 64)
    with pytest.raises(gpu.JpegUnsupported):
        gpu.jpeg_info(data[:30])


def _with_exif_orientation(data, orientation, big_endian=False):
    """Insert an APP1 Exif segment holding only the orientation tag right after SOI."""
    import struct
    e = ">" if big_endian else "<"
    tiff = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, 8) + struct.pack(e + "H", 1) + \
        struct.pack(e + "HHIHH", 0x0112, 3, 1, orientation, 0) + struct.pack(e + "I", 0)
    payload = b"Exif\0\0" + tiff
    return data[:2] + b"\xff\xe1" + struct.pack(">H", len(payload) + 2) + payload + data[2:]


def test_exif_orientation_is_left_to_cv2():
    """cv2.imdecode applies the Exif orientation (the reference never sets IMREAD_IGNORE_ORIENTATION): a tile tagged 2..8 comes out
    rotated / flipped there.  The library decoder refuses such files, so that its callers take their cv2 path (Stitcher._load_sequence)."""
    from imagestitch_b200 import gpu
    img = _image(40, 60, 1)
    data = _encode(img, 90)
    for be in (False, True):
        rot = _with_exif_orientation(data, 6, be)
        assert cv2.imdecode(np.frombuffer(rot, np.uint8), cv2.IMREAD_GRAYSCALE).shape == (60, 40)       # cv2 rotates
        with pytest.raises(gpu.JpegUnsupported):
            gpu.jpeg_info(rot)
        up = _with_exif_orientation(data, 1, be)                                                       # "top-left": nothing to do
        assert gpu.jpeg_info(up)[:2] == (40, 60)
        coef, _ = gpu.jpeg_luma_coefficients(up)
        assert np.array_equal(coef, gpu.jpeg_luma_coefficients(data)[0])


def test_damaged_files_never_crash_the_host_stage():
    """Bit flips, truncations and spliced garbage in headers and entropy data: the parser either reports unsupported / bad input or
    decodes something of the advertised geometry -- no out-of-bounds access (run under the normal allocator; a crash fails the suite)."""
    from imagestitch_b200 import gpu
    rng = np.random.default_rng(99)
    base = [_encode(_image(40, 56, 1), 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, restart=2), _encode(_image(33, 17, 2, channels=1), 70)]
    outcomes = {"ok": 0, "rejected": 0}
    for trial in range(400):
        d = bytearray(base[trial % 2])
        kind = trial % 4
        if kind == 0:                                    # flips anywhere (headers included)
            for _ in range(1 + trial % 5):
                d[int(rng.integers(2, len(d)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:                                  # truncation
            d = d[: int(rng.integers(4, len(d)))]
        elif kind == 2:                                  # garbage run
            p = int(rng.integers(2, len(d) - 8)); d[p:p + 8] = bytes(rng.integers(0, 256, 8, dtype=np.uint8))
        else:                                            # damaged Huffman table counts
            p = bytes(d).find(b"\xff\xc4")
            d[p + 5 + int(rng.integers(0, 16))] = int(rng.integers(0, 256))
        try:
            r, c, n = gpu.jpeg_info(bytes(d))
            if r * c > 1 << 22:
                outcomes["rejected"] += 1                # absurd geometry from a flipped SOF: not decoded here
                continue
            for comp in range(n):
                coef, quant, _ = gpu.jpeg_component_coefficients(bytes(d), comp)
                assert coef.ndim == 3 and coef.shape[2] == 64
            outcomes["ok"] += 1
        except gpu.VfsmsError:
            outcomes["rejected"] += 1
    assert outcomes["ok"] > 50 and outcomes["rejected"] > 50, outcomes
