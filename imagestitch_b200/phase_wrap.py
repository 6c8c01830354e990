"""Wrap-aware phase-correlation offsets -- an opt-in CORRECTED mode beside the parity mode.

The reference's incremental phase search (Stitcher.py:205-258) is wrong by construction: cv2.phaseCorrelate
returns the shift (sx, sy) with  roiB(x, y) = roiA(x - sx, y - sy),  i.e. the ROI-relative offset is (-sy, -sx), yet the code
adds (int(sy), int(sx)); and the shift is only known modulo the padded DFT size (M, N), so overlaps smaller than half the ROI
alias.  `Stitcher.phaseMode = "reference"` (default) keeps that behaviour for parity.  "wrapAware":
  1. phase correlation of the two ROI strips (the same device call),
  2. candidates = (-round(sy) + k M, -round(sx) + l N), k, l in {-1, 0, 1}, that leave at least `min_overlap` of the ROI,
  3. each candidate is scored by the zero-mean normalised cross-correlation of the pixels the two ROIs would share -- integer
     sums (n, Sa, Sb, Sab, Saa, Sbb) from one device reduction over all candidates (vfsms_overlap_sums_host), ZNCC from the
     integers in float64 on the host: exact and order-independent,
  4. the best candidate is accepted when its ZNCC exceeds `accept`.
The two device calls are injected so that the same selection logic runs against the NumPy oracle in the CPU tests.
"""
import numpy as np


def optimal_dft_size(n):
    """Smallest 2^a 3^b 5^c >= n (cv2.getOptimalDFTSize, used by phaseCorrelate to pad)."""
    best = None
    p2 = 1
    while p2 < 2 * n + 2:
        p3 = p2
        while p3 < 2 * n + 2:
            p5 = p3
            while p5 < 2 * n + 2:
                if p5 >= n and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def wrap_candidates(shift, rows, cols, min_overlap=0.05):
    """shift = (sx, sy) as cv2.phaseCorrelate returns it -> ROI-relative offsets (dRow, dCol) with roiB(r, c) = roiA(r + dRow,
    c + dCol), one per alias that leaves an overlap of at least min_overlap * rows * cols pixels."""
    M, N = optimal_dft_size(rows), optimal_dft_size(cols)
    d_row0, d_col0 = -int(np.rint(shift[1])), -int(np.rint(shift[0]))
    out = []
    for k in (0, -1, 1):
        for l in (0, -1, 1):
            dr, dc = d_row0 + k * M, d_col0 + l * N
            if (rows - abs(dr)) > 0 and (cols - abs(dc)) > 0 and (rows - abs(dr)) * (cols - abs(dc)) >= min_overlap * rows * cols:
                out.append((dr, dc))
    return out


def zncc(sums):
    """(n, Sa, Sb, Sab, Saa, Sbb) integers -> zero-mean normalised cross-correlation in [-1, 1] (0 for flat overlaps)."""
    n, sa, sb, sab, saa, sbb = (int(v) for v in sums)
    if n <= 0:
        return 0.0
    cov = n * sab - sa * sb
    va = n * saa - sa * sa
    vb = n * sbb - sb * sb
    if va <= 0 or vb <= 0:
        return 0.0
    return float(cov / (np.sqrt(float(va)) * np.sqrt(float(vb))))


def resolve(roi_a, roi_b, phase_fn, sums_fn, min_overlap=0.05, accept=0.5):
    """-> (status, [dRow, dCol], zncc, response).  phase_fn(a, b) -> ((sx, sy), response);
    sums_fn(a, b, [(dRow, dCol), ...]) -> integer array [n, 6]."""
    rows, cols = roi_a.shape[:2]
    (shift, response) = phase_fn(roi_a, roi_b)
    cands = wrap_candidates(shift, rows, cols, min_overlap)
    if not cands:
        return (False, [0, 0], 0.0, float(response))
    sums = np.asarray(sums_fn(roi_a, roi_b, cands))
    scores = [zncc(s) for s in sums]
    best = int(np.argmax(scores))             # ties -> first candidate (the un-aliased one comes first)
    return (scores[best] > accept, [int(cands[best][0]), int(cands[best][1])], float(scores[best]), float(response))
