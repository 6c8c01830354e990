"""Generates tests/golden/blend_cases.npz and tests/golden/mosaic_case.npz -- run in the build container only.

Drives the UNMODIFIED reference Stitcher.fuseImage / ImageFusion.getWeightsMatrix / Stitcher.getStitchByOffset
(via oracle/reference_shims.py) on seeded random overlap ROIs covering: 1-D ramp vs corner weights, every offset-sign
branch, the three aspect cases, gray and colour, all blend modes.
"""
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shims as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def smooth(rng, shape):
    base = rng.integers(0, 256, shape).astype(np.float32)
    k = 5
    base = cv2.blur(base, (k, k))
    return np.clip(base * 1.6 - 60, 0, 255).astype(np.int64)


def make_case(rng, rows, cols, ch, kind):
    shape = (rows, cols) if ch == 1 else (rows, cols, ch)
    B = smooth(rng, shape)
    A = smooth(rng, shape)
    if kind == "full":
        pass
    elif kind == "mostly":            # > 65 % valid: still the 1-D ramp
        A[: rows // 4, : cols // 3] = -1
    else:                             # corner: only one L-shaped part of A holds data
        A[...] = -1
        r = rng.integers(rows // 4, rows // 2); c = rng.integers(cols // 4, cols // 2)
        Afull = smooth(rng, shape)
        if kind == "corner_ul":       # data in the upper-left part
            A[:r, :] = Afull[:r, :]; A[:, :c] = Afull[:, :c]
        elif kind == "corner_ll":
            A[rows - r:, :] = Afull[rows - r:, :]; A[:, :c] = Afull[:, :c]
        elif kind == "corner_lr":
            A[rows - r:, :] = Afull[rows - r:, :]; A[:, cols - c:] = Afull[:, cols - c:]
        elif kind == "corner_ur":
            A[:r, :] = Afull[:r, :]; A[:, cols - c:] = Afull[:, cols - c:]
    return A, B


def main():
    S, U, F = rs.import_reference()
    rng = np.random.default_rng(20260925)
    cases = {}
    idx = 0
    st = S.Stitcher()
    for ch in (1, 3):
        for (rows, cols) in ((36, 60), (60, 36), (40, 40), (23, 57)):
            for kind in ("full", "mostly", "corner_ul", "corner_ll", "corner_lr", "corner_ur"):
                for (dx, dy) in ((5, 7), (-5, -7), (0, 0), (9, -3), (-2, 4)):
                    if kind.startswith("corner") and (dx, dy) not in ((5, 7), (-5, -7)):
                        continue
                    A, B = make_case(rng, rows, cols, ch, kind)
                    entry = {"A": A.astype(np.int16), "B": B.astype(np.int16), "dx": dx, "dy": dy}
                    for method in ("average", "maximum", "minimum", "fadeInAndFadeOut", "trigonometric", "multiBandBlending"):
                        if method == "multiBandBlending" and (ch == 3 or (dx, dy) != (5, 7)):
                            continue
                        S.Stitcher.fuseMethod = method
                        S.Stitcher.isColorMode = (ch == 3)
                        try:
                            out = st.fuseImage([A.copy(), B.copy()], dx, dy)
                        except Exception as e:          # ZeroDivisionError / IndexError quirks of getWeightsMatrix
                            entry["err_" + method] = type(e).__name__
                            continue
                        entry["out_" + method] = np.asarray(out).astype(np.uint8)
                    if kind.startswith("corner"):
                        st.imageFusion.isColorMode = (ch == 3)
                        try:
                            wa, wb = st.imageFusion.getWeightsMatrix([A.copy(), B.copy()])
                            entry["wa"] = np.asarray(wa, np.float32); entry["wb"] = np.asarray(wb, np.float32)
                        except Exception as e:
                            entry["err_weights"] = type(e).__name__
                    for k, v in entry.items():
                        cases["c%03d_%s" % (idx, k)] = np.asarray(v)
                    cases["c%03d_meta" % idx] = np.array([rows, cols, ch, dx, dy])
                    cases["c%03d_kind" % idx] = np.array(kind)
                    idx += 1
    cases["n_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(OUT, "blend_cases.npz"), **cases)
    print("blend cases", idx)

    # mini mosaic: 8 synthetic 128x116 tiles on a 2-row serpentine, golden output of the reference getStitchByOffset
    from imagestitch_b200 import synth
    tiles, offs = synth.tile_sequence(seed=31, n_rows=2, n_cols=4, size=128, overlap=28, noise=1.0)
    tiles = np.ascontiguousarray(tiles[:, :, :116])   # non-square tiles, 16 px horizontal overlap
    offs = offs.copy()
    mos = {"tiles": tiles, "offsets": offs}
    with tempfile.TemporaryDirectory() as d:
        files = []
        for k, t in enumerate(tiles):
            f = os.path.join(d, "t%02d.png" % k); cv2.imwrite(f, t); files.append(f)
        for color in (False, True):
            for method in ("notFuse", "average", "fadeInAndFadeOut", "trigonometric"):
                S.Stitcher.fuseMethod = method; S.Stitcher.isColorMode = color
                out = st.getStitchByOffset(files, [list(map(int, o)) for o in offs])
                mos["out_%s_%s" % (method, "color" if color else "gray")] = out
    np.savez_compressed(os.path.join(OUT, "mosaic_case.npz"), **mos)
    print("mosaic", {k: v.shape for k, v in mos.items() if k.startswith("out")})


if __name__ == "__main__":
    main()
