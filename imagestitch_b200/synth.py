"""Seeded synthetic micrograph tiles; ground-truth offsets are known exactly.

canvas(seed, H, W): sum over sigma in {1.5, 3, 6, 12} px of Gaussian-blurred standard-normal fields (each normalised
to unit std, weights 1.0/0.8/0.6/0.4) plus dark discs r ~ U[3, 12] at 20 per Mpx, mapped to mean 124 / std 40, u8.
tile sequence: tile_rows x tile_cols crops on a serpentine grid, nominal step = size - overlap, integer jitter
U[-24, 24] along travel and U[-4, 4] across, times a vignette 1 - 0.15 r^2, plus N(0, 2) noise.
"""
import cv2
import numpy as np


def canvas(seed, H, W):
    acc = np.zeros((H, W), np.float32)
    for k, (sigma, wgt) in enumerate(zip((1.5, 3.0, 6.0, 12.0), (1.0, 0.8, 0.6, 0.4))):
        f = np.random.default_rng(seed + k).standard_normal((H, W), dtype=np.float32)
        f = cv2.GaussianBlur(f, (0, 0), sigma, borderType=cv2.BORDER_REFLECT)
        f /= max(float(f.std()), 1e-6)
        acc += np.float32(wgt) * f
    rng = np.random.default_rng(seed + 100)
    n_disc = int(round(20 * H * W / 1e6))
    ys = rng.integers(0, H, n_disc); xs = rng.integers(0, W, n_disc); rs = rng.uniform(3, 12, n_disc)
    depth = rng.uniform(0.8, 2.0, n_disc)
    for y, x, r, d in zip(ys, xs, rs, depth):
        cv2.circle(acc, (int(x), int(y)), int(round(r)), float(acc[y, x] - d), -1, lineType=cv2.LINE_AA)
    acc = (acc - acc.mean()) / max(float(acc.std()), 1e-6)
    return np.clip(acc * 40.0 + 124.0, 0, 255).astype(np.uint8)


def _vignette(h, w):
    yy = (np.arange(h, dtype=np.float32) - (h - 1) / 2) / (h / 2)
    xx = (np.arange(w, dtype=np.float32) - (w - 1) / 2) / (w / 2)
    r2 = yy[:, None] ** 2 + xx[None, :] ** 2
    return (1.0 - 0.15 * r2 / 2.0).astype(np.float32)


def serpentine_origins(n_rows, n_cols, size, overlap, seed):
    """Top-left corners of the tiles in shooting order and the true offsets between consecutive tiles."""
    rng = np.random.default_rng(seed + 7)
    step = size - overlap
    origins = []
    r, c = 32, 32
    for gr in range(n_rows):
        cols = range(n_cols) if gr % 2 == 0 else range(n_cols - 1, -1, -1)
        for i, gc in enumerate(cols):
            if gr == 0 and i == 0:
                origins.append((r, c)); continue
            if i == 0:      # row turn: move down
                r = r + step + int(rng.integers(-24, 25)); c = c + int(rng.integers(-4, 5))
            else:
                sgn = 1 if gr % 2 == 0 else -1
                c = c + sgn * (step + int(rng.integers(-24, 25))); r = r + int(rng.integers(-4, 5))
            origins.append((r, c))
    origins = np.array(origins, np.int64)
    origins[:, 0] -= origins[:, 0].min() - 8
    origins[:, 1] -= origins[:, 1].min() - 8
    offsets = np.diff(origins, axis=0)
    return origins, offsets


def tile_sequence(seed, n_rows, n_cols, size=2048, overlap=205, noise=2.0):
    """Returns (tiles [n, size, size] u8, true_offsets [n-1, 2] (dRow, dCol))."""
    origins, offsets = serpentine_origins(n_rows, n_cols, size, overlap, seed)
    H = int(origins[:, 0].max()) + size + 8
    W = int(origins[:, 1].max()) + size + 8
    base = canvas(seed, H, W).astype(np.float32)
    vig = _vignette(size, size)
    tiles = np.empty((len(origins), size, size), np.uint8)
    for k, (r, c) in enumerate(origins):
        t = base[r:r + size, c:c + size] * vig
        if noise > 0:
            t = t + np.random.default_rng(seed + 1000 + k).normal(0.0, noise, t.shape).astype(np.float32)
        tiles[k] = np.clip(t + 0.5, 0, 255).astype(np.uint8)
    return tiles, offsets


def pair(seed=1234, size=2048, overlap=205, direction=1):
    """Two tiles overlapping along `direction` (1: B below A, 2: B right of A).  Returns A, B, (dRow, dCol)."""
    if direction == 1:
        tiles, off = tile_sequence(seed, 2, 1, size, overlap)
    else:
        tiles, off = tile_sequence(seed, 1, 2, size, overlap)
    return tiles[0], tiles[1], (int(off[0][0]), int(off[0][1]))


# ---------------------------------------------------------------- torch (device) variant of the same recipe, for bench.py
def canvas_torch(seed, H, W, device):
    """Same recipe as canvas() evaluated with torch on `device` (different RNG stream, same statistics)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    acc = torch.zeros((H, W), dtype=torch.float32, device=device)
    for k, (sigma, wgt) in enumerate(zip((1.5, 3.0, 6.0, 12.0), (1.0, 0.8, 0.6, 0.4))):
        g.manual_seed(seed + k)
        f = torch.randn((1, 1, H, W), generator=g, device=device, dtype=torch.float32)
        r = int(3 * sigma + 0.5)
        x = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
        kern = torch.exp(-0.5 * (x / sigma) ** 2)
        kern = kern / kern.sum()
        f = F.conv2d(F.pad(f, (r, r, 0, 0), mode="reflect"), kern.view(1, 1, 1, -1))
        f = F.conv2d(F.pad(f, (0, 0, r, r), mode="reflect"), kern.view(1, 1, -1, 1))
        f = f[0, 0]
        acc += wgt * f / f.std().clamp_min(1e-6)
    rng = np.random.default_rng(seed + 100)
    n_disc = int(round(20 * H * W / 1e6))
    ys = rng.integers(16, H - 16, n_disc); xs = rng.integers(16, W - 16, n_disc)
    rs = rng.uniform(3, 12, n_disc); depth = rng.uniform(0.8, 2.0, n_disc)
    yy, xx = torch.meshgrid(torch.arange(-13, 14, device=device), torch.arange(-13, 14, device=device), indexing="ij")
    d2 = (yy * yy + xx * xx).float()
    for y, x, r, d in zip(ys, xs, rs, depth):
        patch = acc[y - 13:y + 14, x - 13:x + 14]
        patch -= float(d) * (d2 <= float(r * r)).float()
    acc = (acc - acc.mean()) / acc.std().clamp_min(1e-6)
    return (acc * 40.0 + 124.0).clamp(0, 255)


def pair_batch_torch(seed, n_pairs, size, overlap, device, noise=2.0, canvas_hw=None):
    """n_pairs vertically overlapping tile pairs (direction 1) cropped from one synthetic canvas.
    Returns (tiles_a [P,size,size] u8, tiles_b, true_offsets [P,2] int64 numpy)."""
    import torch
    step = size - overlap
    H, W = canvas_hw if canvas_hw else (size + step + 64, size * 3)
    base = canvas_torch(seed, H, W, device)
    rng = np.random.default_rng(seed + 7)
    vig = torch.from_numpy(_vignette(size, size)).to(device)
    g = torch.Generator(device=device)
    A = torch.empty((n_pairs, size, size), dtype=torch.uint8, device=device)
    B = torch.empty_like(A)
    offs = np.zeros((n_pairs, 2), np.int64)
    for p in range(n_pairs):
        dr = step + int(rng.integers(-24, 25)); dc = int(rng.integers(-4, 5))
        r0 = int(rng.integers(0, H - size - dr)); c0 = int(rng.integers(4, W - size - 4))
        offs[p] = (dr, dc)
        for dst, (r, c), sd in ((A, (r0, c0), 2 * p), (B, (r0 + dr, c0 + dc), 2 * p + 1)):
            t = base[r:r + size, c:c + size] * vig
            if noise > 0:
                g.manual_seed(seed + 1000 + sd)
                t = t + noise * torch.randn(t.shape, generator=g, device=device)
            dst[p] = (t + 0.5).clamp(0, 255).to(torch.uint8)
    return A, B, offs


def sequence_torch(seed, n_rows, n_cols, size, overlap, device, first=0, count=None, noise=2.0):
    """Tiles first .. first + count - 1 of the serpentine sequence (n_rows x n_cols, shooting order) cropped from ONE synthetic canvas
    evaluated on `device`: every rank of a sharded run generates the same canvas and crops its own run of tiles.
    Returns (tiles [count, size, size] u8 on device, true pair offsets of the WHOLE sequence [n - 1, 2] (dRow, dCol) numpy)."""
    import torch
    origins, offsets = serpentine_origins(n_rows, n_cols, size, overlap, seed)
    n = len(origins)
    count = n - first if count is None else count
    H = int(origins[:, 0].max()) + size + 8
    W = int(origins[:, 1].max()) + size + 8
    base = canvas_torch(seed, H, W, device)
    vig = torch.from_numpy(_vignette(size, size)).to(device)
    g = torch.Generator(device=device)
    tiles = torch.empty((count, size, size), dtype=torch.uint8, device=device)
    for j in range(count):
        k = first + j
        r, c = int(origins[k, 0]), int(origins[k, 1])
        t = base[r:r + size, c:c + size] * vig
        if noise > 0:
            g.manual_seed(seed + 1000 + k)
            t = t + noise * torch.randn(t.shape, generator=g, device=device)
        tiles[j] = (t + 0.5).clamp(0, 255).to(torch.uint8)
    del base
    return tiles, offsets
