"""CPU: the JPEG-encode oracle (oracle/jpeg_encode_oracle.py) pinned against cv2.imencode -- the call behind the reference's
cv2.imwrite(..., stitchResult) (Stitcher.py:130-131, :196-197) -- byte for byte."""
import numpy as np
import pytest


def _images():
    from imagestitch_b200 import synth
    rng = np.random.default_rng(1)
    A, _, _ = synth.pair(seed=3, size=512, overlap=60, direction=1)
    col = np.dstack([A, np.roll(A, 7, 0), 255 - np.roll(A, 5, 1)])
    out = []
    # 1x1 ... sizes around the 8 / 16 pixel block and MCU boundaries (dummy luma blocks at the right edge and at the bottom)
    for (h, w) in ((1, 1), (8, 8), (7, 9), (16, 16), (17, 33), (24, 40), (40, 24), (9, 16), (16, 9), (131, 75), (200, 264)):
        out.append(("gray%dx%d" % (h, w), np.ascontiguousarray(A[:h, :w])))
        out.append(("bgr%dx%d" % (h, w), np.ascontiguousarray(col[:h, :w])))
    out.append(("noise_gray", rng.integers(0, 256, (67, 93), dtype=np.uint8)))
    out.append(("noise_bgr", rng.integers(0, 256, (67, 93, 3), dtype=np.uint8)))          # long codes, many 0xFF bytes to stuff
    out.append(("flat", np.full((33, 47, 3), 200, np.uint8)))                            # EOB-only blocks
    out.append(("extreme", (rng.integers(0, 2, (48, 48, 3)) * 255).astype(np.uint8)))    # largest coefficient categories
    return out


@pytest.mark.parametrize("quality", [95, 100, 50, 10])
def test_oracle_bytes_equal_cv2(quality):
    import cv2
    from oracle import jpeg_encode_oracle as jo
    for name, im in _images():
        ref = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, quality])[1].tobytes()
        assert jo.encode(im, quality) == ref, (name, quality)


def test_default_quality_is_cv2s():
    import cv2
    from oracle import jpeg_encode_oracle as jo
    _, im = _images()[-3]
    assert jo.encode(im) == cv2.imencode(".jpg", im)[1].tobytes()


def test_golden_demo_tile(golden_dir):
    """one of the reference's demo micrographs (committed fixture): decode with cv2, re-encode, compare"""
    import os
    import cv2
    from oracle import jpeg_encode_oracle as jo
    names = sorted(n for n in os.listdir(golden_dir) if n.lower().endswith((".jpg", ".png")))
    assert names
    img = cv2.imread(os.path.join(golden_dir, names[0]), cv2.IMREAD_COLOR)[:160, :232]
    assert jo.encode(img) == cv2.imencode(".jpg", img)[1].tobytes()
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert jo.encode(gray) == cv2.imencode(".jpg", gray)[1].tobytes()


def test_tables_match_the_file_cv2_writes():
    """the Annex-K tables restated in the oracle are the ones in cv2's DQT / DHT segments"""
    import cv2
    from oracle import jpeg_encode_oracle as jo
    b = cv2.imencode(".jpg", np.zeros((16, 16, 3), np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 50])[1].tobytes()     # quality 50: unscaled tables
    i, dqt, dht = 2, [], []
    while b[i + 1] != 0xDA:
        m, n = b[i + 1], (b[i + 2] << 8) | b[i + 3]
        if m == 0xDB:
            dqt.append(b[i + 4:i + 2 + n])
        if m == 0xC4:
            dht.append(b[i + 4:i + 2 + n])
        i += 2 + n
    assert [bytes(d[1:]) for d in dqt] == [bytes(int(v) for v in t[jo.ZIGZAG]) for t in (jo.STD_LUMA_Q, jo.STD_CHROMA_Q)]
    expect = [(0x00, jo.DC_LUMA), (0x10, jo.AC_LUMA), (0x01, jo.DC_CHROMA), (0x11, jo.AC_CHROMA)]
    assert [bytes(d) for d in dht] == [bytes([i]) + bytes(bits) + bytes(vals) for i, (bits, vals) in expect]
