#!/usr/bin/env python
"""Secondary measurement: mosaic-sized canvas in HBM -> the bytes cv2.imwrite(".jpg") would write.

    python scripts/bench_encode.py [--rows 8192 --cols 8192 --gray] [--reps 5]

Prints one JSON line: device encode (CUDA events around the whole call's kernels are not available through the C ABI, so wall
clock of the synchronous call, D2H of the compressed stream included), cv2.imencode on one host thread beside it, byte identity,
and the algorithmic HBM bytes of the transform kernel.  Needs a GPU; there is no CPU fallback.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--gray", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import cv2
    import torch
    from imagestitch_b200 import gpu, synth
    dev = torch.device("cuda", 0)
    tile = synth.canvas(1234, 2048, 2048) if hasattr(synth, "canvas") else synth.pair(seed=1234, size=2048, overlap=200, direction=1)[0]
    t = torch.from_numpy(np.ascontiguousarray(tile)).to(dev)
    reps_r, reps_c = -(-a.rows // 2048), -(-a.cols // 2048)
    gray = t.repeat(reps_r, reps_c)[:a.rows, :a.cols].contiguous()
    img = gray if a.gray else torch.stack([gray, torch.roll(gray, 5, 0), 255 - torch.roll(gray, 9, 1)], dim=2).contiguous()
    torch.cuda.synchronize(dev)
    data = gpu.jpeg_encode_dev(img)                       # warm-up
    t0 = time.perf_counter()
    for _ in range(a.reps):
        data = gpu.jpeg_encode_dev(img)
    dt = (time.perf_counter() - t0) / a.reps
    host = img.cpu().numpy()
    t0 = time.perf_counter()
    ref = cv2.imencode(".jpg", host)[1].tobytes()
    dt_c = time.perf_counter() - t0
    mp = a.rows * a.cols / 1e6
    ch = 1 if a.gray else 3
    print(json.dumps({
        "what": "%d x %d %s canvas in HBM -> baseline JPEG (q95%s)" % (a.rows, a.cols, "gray" if a.gray else "BGR", "" if a.gray else ", 4:2:0"),
        "device_encode_mpix_per_s": mp / dt, "device_encode_ms": dt * 1e3, "cv2_imencode_mpix_per_s_1_thread": mp / dt_c, "cv2_imencode_ms": dt_c * 1e3,
        "identical_to_cv2": data == ref, "jpeg_bytes": len(data),
        "transform_kernel_algorithmic_bytes": int(a.rows * a.cols * (ch + (2 if a.gray else 3))),
    }), flush=True)


if __name__ == "__main__":
    main()
