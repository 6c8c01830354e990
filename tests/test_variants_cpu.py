"""CPU checks of the invariants the kernel schedules (include/vfsms.h VFSMS_OPT_*) rely on, restated in NumPy from the kernel
source -- the kernels themselves are compared with each other and with the oracle on the GPU (tests/test_gpu_variants.py)."""
import numpy as np

f32 = np.float32


def test_describe_mode2_interior_predicate_is_conservative():
    """surf_describe.cuh describe_fixed_kernel: when `interior` holds, no sample of the rotated window may need the border path
    (the stacked texture would otherwise return the neighbouring image's rows).  Sample positions restated exactly as the
    kernel forms them: float row chain (start += sin / cos), double column positions."""
    from imagestitch_b200 import synth
    from oracle import surf
    checked = 0
    worst = 1e9
    for seed, (R, C) in [(1234, (409, 768)), (5, (97, 384)), (7, (260, 300))]:
        A, _, _ = synth.pair(seed=seed, size=1024, overlap=110, direction=1)
        img = np.ascontiguousarray(A[:R, :C])
        kps, _ = surf.detect_and_compute(img, 30, 4, 3, True, False, 0)
        ncols1, nrows1 = C - 1, R - 1
        for kp in kps[::3]:
            cx, cy, size, ang = f32(kp[0]), f32(kp[1]), f32(kp[2]), f32(kp[3])
            s = f32(size * f32(1.2) / f32(9.0))
            win = int(f32(21) * s)
            if win > 768:
                continue
            rad = f32(f32(win - 1) * f32(0.7072) + f32(2.0))
            if not (cx - rad >= 1 and cx + rad <= f32(ncols1 - 1) and cy - rad >= 1 and cy + rad <= f32(nrows1 - 1)):
                continue
            dir_rad = f32(ang * f32(np.pi / 180))
            sin_dir, cos_dir = f32(-np.sin(np.float64(dir_rad))), f32(np.cos(np.float64(dir_rad)))
            wo = f32(-f32(win - 1) / 2)
            sx = f32(f32(cx + f32(wo * cos_dir)) + f32(wo * sin_dir))
            sy = f32(f32(cy - f32(wo * sin_dir)) + f32(wo * cos_dir))
            xs, ys = np.empty(win, np.float32), np.empty(win, np.float32)
            for i in range(win):
                xs[i], ys[i] = sx, sy
                sx, sy = f32(sx + sin_dir), f32(sy + cos_dir)
            j = np.arange(win, dtype=np.float64)
            X = xs[:, None].astype(np.float64) + j[None, :] * np.float64(cos_dir)
            Y = ys[:, None].astype(np.float64) - j[None, :] * np.float64(sin_dir)
            # floor_split yields floor(x) or floor(x) - 1 (exact integers): both must index a full 2x2 footprint
            assert X.min() >= 1 and Y.min() >= 1 and np.floor(X).max() < ncols1 and np.floor(Y).max() < nrows1, (kp, win)
            worst = min(worst, X.min(), Y.min(), ncols1 - X.max(), nrows1 - Y.max())
            checked += 1
    assert checked > 300 and worst >= 2.0, (checked, worst)


def test_sort_mode1_bin_ranking_equals_full_order():
    """surf.cu bin_* kernels: rank = (candidates in higher 13-bit response bins) + (members of the own bin that sort before),
    cut at the bin that reaches max_features -- restated in NumPy, must reproduce the KeypointGreater order of the strongest
    max_features candidates (response desc, size desc, octave desc, y desc, x asc)."""
    rng = np.random.default_rng(4)
    for trial in range(20):
        n = int(rng.integers(1, 5000))
        max_features = int(rng.integers(1, 4000)) if trial % 4 else 0
        resp = (100 + rng.gamma(1.2, 900.0, n)).astype(np.float32)
        resp[rng.integers(0, n, n // 10)] = resp[rng.integers(0, n, n // 10)]          # exact ties
        size = rng.integers(9, 200, n).astype(np.float32); octave = rng.integers(0, 4, n).astype(np.float32)
        y = rng.uniform(0, 400, n).astype(np.float32); x = rng.uniform(0, 2000, n).astype(np.float32)
        order = sorted(range(n), key=lambda i: (-resp[i], -size[i], -octave[i], -y[i], x[i]))
        n_keep = min(n, max_features) if max_features else n
        bits = resp.view(np.uint32)
        bins = bits >> 19
        hist = np.bincount(bins, minlength=1 << 13)
        start = np.concatenate([np.cumsum(hist[::-1])[::-1][1:], [0]])          # candidates in higher bins
        thr_bin = 0
        if max_features and n > max_features:
            incl = start + hist
            thr_bin = int(np.max(np.flatnonzero(incl >= max_features)))
        staged = np.flatnonzero(bins >= thr_bin)
        out = {}
        for i in staged:
            same = staged[bins[staged] == bins[i]]
            before = sum(1 for jj in same if jj != i and
                         (-resp[jj], -size[jj], -octave[jj], -y[jj], x[jj]) < (-resp[i], -size[i], -octave[i], -y[i], x[i]))
            rank = int(start[bins[i]]) + before
            if rank < n_keep:
                assert rank not in out
                out[rank] = int(i)
        assert [out[r] for r in range(n_keep)] == order[:n_keep]
