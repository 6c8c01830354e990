"""TEST INFRASTRUCTURE ONLY.  Small inputs through every kernel family of the emulated library built with ThreadSanitizer
(VFSMS_EMU_SANITIZE=thread): every CUDA thread is a TSan fiber, only barriers / warp collectives order them, so two threads of a
block touching the same word without a barrier in between are reported -- racecheck for intra-block races.

    VFSMS_EMU_SANITIZE=thread python tests/cuda_emu/build_emu.py
    LD_PRELOAD=$(/usr/bin/g++ -print-file-name=libtsan.so) TSAN_OPTIONS="suppressions=tests/cuda_emu/tsan.supp exitcode=0" \\
        python tests/cuda_emu/racecheck.py 2> racecheck.log ; grep -c "WARNING: ThreadSanitizer" racecheck.log
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    import cv2
    os.environ["VFSMS_EMU_SANITIZE"] = "thread"
    import build_emu
    from imagestitch_b200 import _lib
    _lib.SO_PATH = build_emu.build()
    from imagestitch_b200 import gpu, synth
    h = _lib.context(0)
    _lib.load().vfsms_set_matcher(h, 1)                    # exact SIMT matcher (match_tc.cu is not emulated)
    only = set(sys.argv[1:])

    def step(name, fn):
        if only and name not in only:
            return
        t = time.time()
        fn()
        print("racecheck %-10s %.1f s" % (name, time.time() - t), flush=True)

    A, B, _ = synth.pair(seed=5, size=384, overlap=60, direction=1)
    roiA, roiB = np.ascontiguousarray(A[384 - 76:, :]), np.ascontiguousarray(B[:76, :])

    def surf():
        for mode in (1, 2, 3, 0):
            gpu.set_option("describe", mode)
            gpu.surf_detect_and_describe(roiA, hessian_threshold=100.0, extended=True, keypoints_ratio=0.01)
        gpu.set_option("describe", 1)
        gpu.set_option("sort", 1); gpu.surf_detect_and_describe(roiA, extended=False, keypoints_ratio=0.0); gpu.set_option("sort", 0)
        gpu.set_option("lpt", 1); gpu.surf_detect_and_describe(roiB, hessian_threshold=30.0); gpu.set_option("lpt", 0)

    def align():
        gpu.align_batch(np.stack([roiA, roiA]), np.stack([roiB, roiB]))

    def match():
        rng = np.random.default_rng(1)
        a = rng.standard_normal((130, 64)).astype(np.float32); b = rng.standard_normal((97, 64)).astype(np.float32)
        m = gpu.match_descriptors(a, b, 2, 0.9)
        k = rng.uniform(0, 100, (130, 2)).astype(np.float32)
        gpu.offset_by_mode(k, k[:97] + 3, m, 1)
        gpu.match_descriptors(rng.integers(0, 255, (70, 32)).astype(np.float32), rng.integers(0, 255, (50, 32)).astype(np.float32), 3, 80.0)

    def blend():
        rng = np.random.default_rng(2)
        a = rng.integers(0, 255, (40, 72)).astype(np.int16); b = rng.integers(0, 255, (40, 72)).astype(np.int16)
        a[:10, :30] = -1
        for method in ("average", "fadeInAndFadeOut", "trigonometric", "multiBandBlending"):
            gpu.fuse_roi(a, b, method, 5, -3)
        tiles = rng.integers(0, 255, (3, 32, 48), dtype=np.uint8)
        from imagestitch_b200 import sharding as sh
        full = [[0, 0], [20, 4], [-3, 30]]
        origins, rois, shape = sh.rectify_offsets(full, [(32, 48)] * 3)
        gpu.mosaic(tiles, np.asarray(origins, np.int32), rois, np.asarray(full, np.int32), "fadeInAndFadeOut", shape)

    def phase():
        gpu.phase_correlate(roiA[:32, :48], roiB[:32, :48])
        gpu.overlap_sums(roiA[:32, :48], roiB[:32, :48], [(3, -2), (0, 0)])

    def enhance():
        gpu.enhance(roiA, clahe=True, clip_limit=20.0, tile_size=5)
        gpu.enhance(roiA, clahe=False, clip_limit=20.0, tile_size=5)

    def orb():
        gpu.orb_detect_and_describe(A[:96, :128])

    col = np.dstack([A[:40, :56], A[40:80, :56], A[80:120, :56]])

    def encode():
        gpu.jpeg_encode(col); gpu.jpeg_encode(np.ascontiguousarray(A[:33, :47]), 50)

    def decode():
        data = cv2.imencode(".jpg", col, [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
        for mode in (0, 1):
            gpu.set_option("entropy", mode)
            gpu.jpeg_decode_gray(data); gpu.jpeg_decode_bgr(data)
        gpu.set_option("entropy", 0)

    for name, fn in (("surf", surf), ("align", align), ("match", match), ("blend", blend), ("phase", phase), ("enhance", enhance), ("orb", orb),
                     ("encode", encode), ("decode", decode)):
        step(name, fn)
    print("racecheck done", flush=True)


if __name__ == "__main__":
    main()
