"""Build container only (marker `reference`): the sequencing layer -- pair loop, segment restart on a failed pair, trailing
single tile, offset rectification, paste/blend -- of our Stitcher against the UNMODIFIED reference Stitcher
(Stitcher.py:49-182, 369-486) driven by the same scripted offset callback.  No GPU: tiles are decoded by cv2 and the device
mosaic is replaced by the NumPy oracle renderer (the CUDA mosaic itself is pinned to the same reference outputs in
tests/test_gpu_blend.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.reference


def _write_tiles(tmp_path, n, h, w, seed):
    import cv2
    rng = np.random.default_rng(seed)
    files = []
    for k in range(n):
        t = rng.integers(1, 255, (h, w)).astype(np.uint8)
        t = cv2.blur(t, (3, 3))
        f = str(tmp_path / ("t%03d.png" % k))
        cv2.imwrite(f, t)
        files.append(f)
    return files


@pytest.mark.parametrize("fuse", ["notFuse", "fadeInAndFadeOut", "average"])
@pytest.mark.parametrize("fail_at", [(), (3,), (0,), (5,), (2, 3), (1, 4)])
def test_flow_stitch_with_multiple_equals_reference(tmp_path, monkeypatch, fuse, fail_at):
    from oracle import reference_shims as rs
    from oracle import blend_oracle as bo
    from imagestitch_b200 import gpu
    from imagestitch_b200.Stitcher import Stitcher
    S, U, F = rs.import_reference()
    n, h, w = 7, 40, 48
    files = _write_tiles(tmp_path, n, h, w, 3)
    offsets = [[2, w - 9], [-1, w - 11], [h - 8, -3], [1, -(w - 10)], [-2, -(w - 12)], [h - 9, 2]]

    # the reference: a failed pair is not retried, the next segment starts after it (Stitcher.py:106-126)
    ref = S.Stitcher()
    monkeypatch.setattr(S.Stitcher, "isPrintLog", False, raising=False)
    monkeypatch.setattr(S.Stitcher, "isColorMode", False, raising=False)
    monkeypatch.setattr(S.Stitcher, "fuseMethod", fuse, raising=False)
    pair_of_call = []
    # scripted offset 'method' keyed by the images themselves (robust against either implementation's visiting order): pair k of
    # the whole sequence gets offsets[k], pairs in fail_at fail
    import cv2
    decoded = [cv2.imread(f, 0) for f in files]

    def method(images):
        a = next(i for i, t in enumerate(decoded) if t.shape == images[0].shape and np.array_equal(t, images[0]))
        b = next(i for i, t in enumerate(decoded) if t.shape == images[1].shape and np.array_equal(t, images[1]))
        assert b == a + 1
        pair_of_call.append(a)
        if a in fail_at:
            return (False, "  The two images can not match")
        return (True, list(offsets[a]))
    out_ref = ref.flowStitchWithMutiple(list(files), method)
    visited_ref = list(pair_of_call)
    pair_of_call.clear()

    def fake_mosaic(tiles, tile_origin, roi_rect, pair_offset, method_name, canvas_shape, device=0):
        out, _ = bo.band_renderer()(tiles, tile_origin, roi_rect, pair_offset, method_name, canvas_shape, False, None, None, None)
        return out
    monkeypatch.setattr(gpu, "mosaic", fake_mosaic)
    ours = Stitcher()
    monkeypatch.setattr(Stitcher, "isPrintLog", False)
    monkeypatch.setattr(Stitcher, "isColorMode", False)
    monkeypatch.setattr(Stitcher, "fuseMethod", fuse)
    monkeypatch.setattr(Stitcher, "decoder", "cv2")
    out_ours = ours.flowStitchWithMutiple(list(files), method)
    assert pair_of_call == visited_ref                       # same pairs visited in the same order
    assert len(out_ours) == len(out_ref)
    for a, b in zip(out_ours, out_ref):
        assert np.asarray(a).shape == np.asarray(b).shape and np.array_equal(a, b)
