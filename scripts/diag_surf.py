import numpy as np, sys
sys.path.insert(0, '.')
from imagestitch_b200 import gpu, synth
from oracle import surf
A, B, off = synth.pair(seed=77, size=1024, overlap=110, direction=1)
L = 204
img = A[1024 - L:]
mf = int(0.01 * img.size)
for ratio, ext in ((0.01, True), (0.0, False)):
    kp_o, d_o = surf.detect_and_compute(img, 100, 4, 3, ext, False, int(ratio * img.size) if ratio > 0 else 0)
    kp_g, d_g = gpu.surf_detect_and_describe(img, hessian_threshold=100, extended=ext, keypoints_ratio=ratio)
    print("ratio", ratio, "n", len(kp_o), len(kp_g))
    n = min(len(kp_o), len(kp_g))
    for col in range(7):
        neq = np.nonzero(kp_o[:n, col] != kp_g[:n, col])[0]
        print(" col", col, "mismatch", len(neq), neq[:5], kp_o[neq[:3], col] if len(neq) else "", kp_g[neq[:3], col] if len(neq) else "")
    so = {tuple(r[[0, 1, 2, 4]]) for r in kp_o}; sg = {tuple(r[[0, 1, 2, 4]]) for r in kp_g}
    print(" set diff", len(so - sg), len(sg - so))
