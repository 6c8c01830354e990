"""NumPy restatements of the non-SURF pieces of the hot path -- TEST INFRASTRUCTURE ONLY.

Each function cites the reference call it follows and is pinned by tests/test_oracle_pins.py against cv2 (available in
this image and on the GPU box) or against fixtures generated from the unmodified reference.
"""
import numpy as np


def optimal_dft_size(n):
    """cv2.getOptimalDFTSize: smallest 2^a 3^b 5^c >= n (phaseCorrelate pads to it; Stitcher.py:230)."""
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


def phase_correlate(a, b):
    """cv2.phaseCorrelate(np.float64(a), np.float64(b)) without window (Stitcher.py:230).
    -> ((shift_x, shift_y), response)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    rows, cols = a.shape
    M, N = optimal_dft_size(rows), optimal_dft_size(cols)
    pa = np.zeros((M, N)); pb = np.zeros((M, N))
    pa[:rows, :cols] = a; pb[:rows, :cols] = b
    Fa, Fb = np.fft.fft2(pa), np.fft.fft2(pb)
    P = Fa * np.conj(Fb)
    mag = np.abs(P)
    C = P * mag / (mag * mag + np.finfo(np.float64).eps)       # divSpectrums(P, |P|) with its eps guard
    c = np.real(np.fft.ifft2(C)) * (M * N)                      # cv2.idft without DFT_SCALE
    c = np.fft.fftshift(c)
    py, px = np.unravel_index(np.argmax(c), c.shape)            # minMaxLoc: first maximum in raster order
    r0, r1 = max(py - 2, 0), min(py + 2, M - 1)
    c0, c1 = max(px - 2, 0), min(px + 2, N - 1)
    win = c[r0:r1 + 1, c0:c1 + 1]
    ys, xs = np.mgrid[r0:r1 + 1, c0:c1 + 1]
    s = win.sum()
    tx = (xs * win).sum() / (s + np.finfo(np.float64).eps)
    ty = (ys * win).sum() / (s + np.finfo(np.float64).eps)
    return (N / 2.0 - tx, M / 2.0 - ty), s / (M * N)


def roi_for_incre(image, direction, order, ratio):
    """Method.getROIRegionForIncreMethod (ImageUtility.py:66-101)."""
    row, col = image.shape[:2]
    if direction in (1, 3):
        n = int(np.floor(row * ratio))
        bottom = (direction == 1) == (order == "first")
        return image[row - n:row, :] if bottom else image[0:n, :]
    n = int(np.floor(col * ratio))
    right = (direction == 2) == (order == "first")
    return image[:, col - n:col] if right else image[:, 0:n]


def overlap_sums(a, b, shifts):
    """Integer sums over the pixels two equally sized ROIs share under roiB(r, c) = roiA(r + dRow, c + dCol), per candidate
    shift: (n, Sa, Sb, Sab, Saa, Sbb).  Oracle of vfsms_overlap_sums_host (the scoring step of the wrap-aware phase mode,
    imagestitch_b200/phase_wrap.py; not part of the reference, which has no such check)."""
    a = np.asarray(a, np.int64); b = np.asarray(b, np.int64)
    rows, cols = a.shape
    out = np.zeros((len(shifts), 6), np.int64)
    for k, (dr, dc) in enumerate(shifts):
        r0, r1 = max(0, -dr), min(rows, rows - dr)
        c0, c1 = max(0, -dc), min(cols, cols - dc)
        if r1 <= r0 or c1 <= c0:
            continue
        pb = b[r0:r1, c0:c1]
        pa = a[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
        out[k] = (pa.size, pa.sum(), pb.sum(), (pa * pb).sum(), (pa * pa).sum(), (pb * pb).sum())
    return out
