"""CPU restatement of the reference's overlap blending and mosaic assembly -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(imagestitch_b200/) never does.  NumPy, float32 / float64 exactly where the reference uses them.

What it follows (reference file:line, tree /root/reference):
  fuse_image            Stitcher.fuseImage                      Stitcher.py:488-525
  fuse_average/max/min  ImageFusion.fuseByAverage/Maximum/...   ImageFusion.py:12-41
  ramp_weights          the ratio > 0.65 branches of fuseByFadeInAndFadeOut / fuseByTrigonometric
                                                                ImageFusion.py:209-235, :261-281
  corner_weights        ImageFusion.getWeightsMatrix            ImageFusion.py:43-190
  fuse_fade / fuse_trig                                         ImageFusion.py:192-244, :246-293
  rectify               Stitcher.getStitchByOffset bookkeeping  Stitcher.py:386-431 (closed form, see below)
  rectify_literal       the same, loop by loop                  Stitcher.py:386-431
  mosaic                Stitcher.getStitchByOffset paste loop   Stitcher.py:433-486

Pinned (tests/test_oracle_pins.py) against tests/golden/blend_cases.npz and mosaic_case.npz, which were produced by the
UNMODIFIED reference (oracle/make_golden_blend.py): weight matrices and average / maximum / minimum / fade / trigonometric
outputs bit-exact, including the cases where the reference raises (IndexError / ZeroDivisionError are re-raised here).
Multi-band blending is not restated (cv2.pyrDown / pyrUp are the oracle for it, see tests/test_gpu_blend.py).
"""
import math

import numpy as np


def _filled(A, color):
    """Per-pixel 'holds data' test of getWeightsMatrix: gray `!= -1`, colour `sum != -3` (ImageFusion.py:73, ...)."""
    return (A.sum(axis=2) != -3) if color else (A != -1)


# One row per corner case of getWeightsMatrix, keyed by the argmin of the quadrant counts [UL, LL, LR, UR]
# (ImageFusion.py:57-62).  columns: order in which columns are scanned; rows_up: rows scanned bottom -> top (hit gives i + 1)
# or top -> bottom (hit gives i - 1); cols_from_right: the second scan walks row `rowIndex` right -> left (hit gives i + 1) or
# left -> right (hit gives i - 1).  The ramps follow from the same two flags.
_CASES = {
    2: dict(columns="right", rows_up=True, cols_from_right=True),     # ImageFusion.py:63-91
    3: dict(columns="right", rows_up=False, cols_from_right=True),    # :93-122
    0: dict(columns="left", rows_up=False, cols_from_right=False),    # :124-153
    1: dict(columns="left", rows_up=True, cols_from_right=False),     # :155-186
}


def corner_weights(A, color=False):
    """ImageFusion.getWeightsMatrix -> (weightA, weightB) float32, shape of A.  Raises what the reference raises."""
    A = np.asarray(A)
    row, col = A.shape[:2]
    quads = [A[0:row // 2, 0:col // 2], A[row // 2:row, 0:col // 2], A[row // 2:row, col // 2:col], A[0:row // 2, col // 2:col]]
    counts = [int(np.count_nonzero(q > 0)) for q in quads]
    case = _CASES[counts.index(min(counts))]
    filled = _filled(A, color)

    # scan 1: walk columns until one yields a non-zero rowIndex (the reference's `if rowIndex != 0: break`)
    col_order = range(col - 1, 0, -1) if case["columns"] == "right" else range(0, col)
    row_index = 0
    for c in col_order:
        hits = np.flatnonzero(filled[:, c])
        if hits.size:
            row_index = int(hits[-1]) + 1 if case["rows_up"] else int(hits[0]) - 1
        if row_index != 0:
            break
    # scan 2: first filled pixel of row `rowIndex` (negative rowIndex wraps like Python indexing; rowIndex == row raises)
    line = filled[row_index]
    hits = np.flatnonzero(line)
    col_index = 0
    if hits.size:
        col_index = int(hits[-1]) + 1 if case["cols_from_right"] else int(hits[0]) - 1

    w_rows = np.ones(A.shape, np.float32)
    w_cols = np.ones(A.shape, np.float32)
    if case["rows_up"]:                        # rows 0 .. rowIndex ramp up to 1: k / rowIndex
        for i in range(row_index + 1):
            if row_index == 0:
                row_index = 1
            w_rows[row_index - i, :] = (row_index - i) * 1 / row_index
    else:                                      # rows rowIndex .. row-1 ramp down to 0
        for i in range(row_index, row):
            if row_index == 0:
                row_index = 1
            w_rows[i, :] = (row - i - 1) * 1 / (row - row_index - 1)
    if case["cols_from_right"]:
        for i in range(col_index + 1):
            if col_index == 0:
                col_index = 1
            w_cols[:, col_index - i] = (col_index - i) * 1 / col_index
    else:
        for i in range(col_index, col):
            if col_index == 0:
                col_index = 1
            w_cols[:, i] = (col - i - 1) * 1 / (col - col_index - 1)
    wB = w_rows * w_cols
    return 1 - wB, wB


def ramp_weights(shape, dx, dy, dtype, trig_layout):
    """The 1-D ramps of the `ratio > 0.65` branch.  fade (float32, ImageFusion.py:213-235) and trigonometric (float64,
    :262-281) index the two matrices differently; `trig_layout` selects which."""
    row, col = shape[:2]
    wA = np.ones(shape, dtype)
    wB = np.ones(shape, dtype)
    if col <= row:                              # horizontal seam: ramp over columns, sign of dy (column offset)
        for i in range(col):
            v = i if dy >= 0 else (col - i)
            a_idx, b_idx = (i, col - i - 1) if trig_layout else (col - i - 1, i)
            wA[:, a_idx] = wA[:, a_idx] * v * 1.0 / col
            wB[:, b_idx] = wB[:, b_idx] * v * 1.0 / col
    else:                                       # vertical seam: ramp over rows, sign of dx (row offset)
        for i in range(row):
            v = i if dx <= 0 else (row - i)
            wA[i, :] = wA[i, :] * v * 1.0 / row
            wB[row - i - 1, :] = wB[row - i - 1, :] * v * 1.0 / row
    return wA, wB


def _weights(A, dx, dy, color, dtype, trig_layout):
    if np.count_nonzero(A > -1) / A.size > 0.65:
        return ramp_weights(A.shape, dx, dy, dtype, trig_layout)
    return corner_weights(A, color)


def _weighted(A, B, wA, wB):
    A = A.copy()
    A[A < 0] = B[A < 0]
    res = wA * A.astype(np.int64) + wB * B.astype(np.int64)
    res[res < 0] = 0
    res[res > 255] = 255
    return res.astype(np.uint8)                 # truncation, like np.uint8(result)


def fuse_fade(A, B, dx, dy, color=False):
    wA, wB = _weights(A, dx, dy, color, np.float32, False)
    return _weighted(A, B, wA, wB)


def fuse_trig(A, B, dx, dy, color=False):
    wA, _ = _weights(A, dx, dy, color, np.float64, True)
    wA = np.power(np.sin(wA * math.pi / 2), 2)
    return _weighted(A, B, wA, 1 - wA)


def fuse_image(A, B, method, dx=0, dy=0, color=False):
    """Stitcher.fuseImage: A, B integer arrays of one overlap ROI (-1 = empty) -> uint8-valued array."""
    A = np.array(A, np.int64)
    B = np.array(B, np.int64)
    if method not in ("fadeInAndFadeOut", "trigonometric"):
        A[A == -1] = 0
        B[B == -1] = 0
        A[A == 0] = B[A == 0]
        B[B == 0] = A[B == 0]
    if method == "notFuse":
        return B
    if method == "average":
        return ((A + B) / 2).astype(np.uint8)
    if method == "maximum":
        return np.maximum(A, B)
    if method == "minimum":
        return np.minimum(A, B)
    if method == "fadeInAndFadeOut":
        return fuse_fade(A, B, dx, dy, color)
    if method == "trigonometric":
        return fuse_trig(A, B, dx, dy, color)
    raise ValueError("method %r is not restated here" % (method,))


def rectify(pair_offsets, tile_shape):
    """The origin shifting loop of getStitchByOffset in closed form.  With cum_i the running sum of the pair offsets
    (cum_0 = 0) the loop keeps the canvas origin at the running minimum, so in final coordinates
        origin_i = cum_i - gmin,  occupied_lo_i = min(0, min_{k<=i} cum_k) - gmin,  occupied_hi_i = max_{k<=i}(cum_k + size) - gmin
    and ROI_i = [max(origin_i, lo_{i-1}), min(origin_i + size, hi_{i-1})) per axis (Stitcher.py:466-469).
    pair_offsets: [n-1, 2] (dRow, dCol).  -> (origins [n, 2], rois [n, 4], (rows, cols))."""
    off = np.concatenate([np.zeros((1, 2), np.int64), np.asarray(pair_offsets, np.int64).reshape(-1, 2)])
    size = np.asarray(tile_shape[:2], np.int64)
    cum = np.cumsum(off, axis=0)
    gmin = np.minimum(cum.min(axis=0), 0)
    origins = cum - gmin
    lo = np.minimum(np.minimum.accumulate(cum, axis=0), 0) - gmin
    hi = np.maximum.accumulate(cum + size, axis=0) - gmin
    rois = np.zeros((len(off), 4), np.int64)
    rois[1:, 0:2] = np.maximum(origins[1:], lo[:-1])
    rois[1:, 2:4] = np.minimum(origins[1:] + size, hi[:-1])
    return origins, rois, (int(hi[-1, 0]), int(hi[-1, 1]))


def rectify_literal(origin_offsets, tile_shapes):
    """Global offset rectification, Stitcher.getStitchByOffset's integer bookkeeping, restated loop by loop (Stitcher.py:386-431:
    the running sums, the O(n^2) shifting of everything placed so far, the per-tile `range` records) -- the checker of the
    product's single-pass form (imagestitch_b200.sharding.rectify_offsets) and of the closed form `rectify` above.
    origin_offsets: [[dRow, dCol], ...] per tile, entry 0 = [0, 0]; tile_shapes: [(rows, cols), ...].
    -> (origins int32 [n, 2], rois int32 [n, 4] (r0, c0, r1, c1; row 0 unused), (canvas_rows, canvas_cols))."""
    n = len(origin_offsets)
    off = [[int(o[0]), int(o[1])] for o in origin_offsets]
    range_x = [[0, 0] for _ in range(n)]
    range_y = [[0, 0] for _ in range(n)]
    result_row, result_col = int(tile_shapes[0][0]), int(tile_shapes[0][1])
    range_x[0][1], range_y[0][1] = result_row, result_col
    dx_sum = dy_sum = 0
    for i in range(1, n):
        h, w = int(tile_shapes[i][0]), int(tile_shapes[i][1])
        dx_sum += off[i][0]
        dy_sum += off[i][1]
        if dx_sum <= 0:
            for j in range(i):
                off[j][0] += abs(dx_sum)
                range_x[j][0] += abs(dx_sum)
                range_x[j][1] += abs(dx_sum)
            result_row += abs(dx_sum)
            range_x[i][1] = result_row
            dx_sum = range_x[i][0] = off[i][0] = 0
        else:
            off[i][0] = dx_sum
            result_row = max(result_row, dx_sum + h)
            range_x[i][1] = result_row
        if dy_sum <= 0:
            for j in range(i):
                off[j][1] += abs(dy_sum)
                range_y[j][0] += abs(dy_sum)
                range_y[j][1] += abs(dy_sum)
            result_col += abs(dy_sum)
            range_y[i][1] = result_col
            dy_sum = range_y[i][0] = off[i][1] = 0
        else:
            off[i][1] = dy_sum
            result_col = max(result_col, dy_sum + w)
            range_y[i][1] = result_col
    rois = np.zeros((n, 4), np.int32)
    for i in range(1, n):
        h, w = int(tile_shapes[i][0]), int(tile_shapes[i][1])
        rois[i] = (max(off[i][0], range_x[i - 1][0]), max(off[i][1], range_y[i - 1][0]),
                   min(off[i][0] + h, range_x[i - 1][1]), min(off[i][1] + w, range_y[i - 1][1]))
    return np.asarray(off, np.int32).reshape(n, 2), rois, (result_row, result_col)


def paste_blend(canvas, tile, origin, roi, pair_offset, method, color, fuse):
    """One iteration of the paste loop on an int canvas with -1 = empty (Stitcher.py:440-483)."""
    h, w = tile.shape[:2]
    r0, c0 = int(origin[0]), int(origin[1])
    if not fuse or method == "notFuse":
        canvas[r0:r0 + h, c0:c0 + w] = tile
        return
    a0, b0, a1, b1 = (int(v) for v in roi)
    A = canvas[a0:a1, b0:b1].copy()
    canvas[r0:r0 + h, c0:c0 + w] = tile
    B = canvas[a0:a1, b0:b1].copy()
    canvas[a0:a1, b0:b1] = fuse_image(A, B, method, int(pair_offset[0]), int(pair_offset[1]), color)


def mosaic(tiles, pair_offsets, method):
    """Stitcher.getStitchByOffset on decoded tiles [n, h, w] or [n, h, w, 3] -> uint8 mosaic."""
    tiles = np.asarray(tiles)
    color = tiles.ndim == 4
    origins, rois, shape = rectify(pair_offsets, tiles.shape[1:3])
    off = np.concatenate([np.zeros((1, 2), np.int64), np.asarray(pair_offsets, np.int64).reshape(-1, 2)])
    canvas = np.zeros(shape + ((3,) if color else ()), np.int64) - 1
    for i in range(len(tiles)):
        paste_blend(canvas, tiles[i], origins[i], rois[i], off[i], method, color, fuse=i > 0)
    canvas[canvas == -1] = 0
    return canvas.astype(np.uint8)


def band_renderer():
    """The `render` callable of imagestitch_b200.sharding.mosaic_sharded, on the CPU (tests of the band partition)."""
    def render(tiles, origins, rois, pair_offsets, method, shape, fuse_first, halo_in, halo_in_rect, halo_out_rect):
        tiles = np.asarray(tiles)
        color = tiles.ndim == 4
        canvas = np.zeros(tuple(shape) + ((3,) if color else ()), np.int64) - 1
        if halo_in is not None:
            r0, c0, hr, hc = (int(v) for v in halo_in_rect)
            canvas[r0:r0 + hr, c0:c0 + hc] = halo_in
        for i in range(len(tiles)):
            paste_blend(canvas, tiles[i], origins[i], rois[i], pair_offsets[i], method, color, fuse=(i > 0 or fuse_first))
        halo_out = None
        if halo_out_rect is not None:
            r0, c0, hr, hc = (int(v) for v in halo_out_rect)
            halo_out = canvas[r0:r0 + hr, c0:c0 + hc].astype(np.int16)
        out = canvas.copy()
        out[out == -1] = 0
        return out.astype(np.uint8), halo_out
    return render
