#!/usr/bin/env python
"""Secondary measurement (not the driver's bench line): the WHOLE reference workflow, files in -> file out, as Main.py runs it.

    python scripts/bench_sequence.py --rows 3 --cols 4                  # 12 synthetic 2048^2 JPEG tiles, colour mode, fade blend, .jpg result

A synthetic serpentine tile grid (SURVEY.md 8(d) generator, the shape of BASELINE.json configs[3] / [4]) is written as JPEG files;
then `Stitcher.imageSetStitchWithMutiple(dir, out, 1, st.calculateOffsetForFeatureSearchIncre, fileExtension="jpg",
outputfileExtension="jpg")` runs with Main.py's settings (SURF, isColorMode = True, fadeInAndFadeOut, roiRatio 0.2, directIncre 1):
tile decode into the HBM stack (gray + colour twin), ROI search over all pairs, colour mosaic with fade blend, JPEG encode of the
result.  Reported: tiles/s of the whole call (wall clock, after one warm-up call), the split decode / align / mosaic / encode, the
offsets against the generator's ground truth, and -- `--cpu-pairs N` -- the same stages on the host: cv2.imdecode (gray + colour per
tile, as the reference does), the oracle's C SURF + cv2 BFMatcher + vote port per pair, the NumPy restatement of the reference's
paste / blend loop, cv2.imwrite; bounded sample, kind "port" (cv2 here has no SURF).  Needs a GPU; there is no CPU fallback.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=3)
    ap.add_argument("--cols", type=int, default=4)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--overlap", type=int, default=205)
    ap.add_argument("--gray", action="store_true", help="isColorMode = False")
    ap.add_argument("--fuse", default="fadeInAndFadeOut")
    ap.add_argument("--options", default="", help='kernel variants, e.g. "entropy=0,describe=0"')
    ap.add_argument("--cpu-pairs", type=int, default=2, help="pairs the host-side port is timed on (0: skip)")
    ap.add_argument("--keep", action="store_true", help="keep the temporary directory")
    a = ap.parse_args()
    import cv2
    from imagestitch_b200 import gpu, synth
    from imagestitch_b200 import Stitcher as S

    tiles, offs = synth.tile_sequence(seed=2025, n_rows=a.rows, n_cols=a.cols, size=a.size, overlap=a.overlap)
    n = len(tiles)
    root = tempfile.mkdtemp(prefix="vfsms_seq_")
    try:
        d = os.path.join(root, "set", "1")
        os.makedirs(d)
        for k, t in enumerate(tiles):
            img = t if a.gray else np.stack([t, np.roll(t, 2, axis=1), 255 - t // 2], axis=-1)
            cv2.imwrite(os.path.join(d, "tile-%04d.jpg" % k), img, [cv2.IMWRITE_JPEG_QUALITY, 92])
        for item in filter(None, a.options.split(",")):
            name, v = item.split("=")
            gpu.set_option(name, int(v))

        St = S.Stitcher
        St.featureMethod = "surf"; St.isColorMode = not a.gray; St.isGPUAvailable = False; St.isEnhance = False
        St.searchRatio = 0.75; St.offsetCaculate = "mode"; St.offsetEvaluate = 3; St.roiRatio = 0.2
        St.fuseMethod = a.fuse; St.isPrintLog = False
        split = {"decode": 0.0, "mosaic": 0.0, "encode": 0.0}
        seen = {}

        def timed(mod, name, key):
            f = getattr(mod, name)

            def g(*args, **kw):
                t0 = time.perf_counter()
                try:
                    return f(*args, **kw)
                finally:
                    split[key] += time.perf_counter() - t0
            setattr(mod, name, g)
            return f
        orig_load = timed(S, "_load_sequence", "decode")
        orig_write = timed(S, "_imwrite", "encode")

        def run(out_dir):
            St.direction = 1; St.directIncre = 1
            st = St()
            orig = st.getStitchByOffset

            def spy(fileList, offsetList):
                seen["offsets"] = [list(o) for o in offsetList]
                t0 = time.perf_counter()
                try:
                    return orig(fileList, offsetList)
                finally:
                    split["mosaic"] += time.perf_counter() - t0
            st.getStitchByOffset = spy
            for k in split:
                split[k] = 0.0
            t0 = time.perf_counter()
            st.imageSetStitchWithMutiple(os.path.join(root, "set"), out_dir, 1, st.calculateOffsetForFeatureSearchIncre, fileExtension="jpg",
                                         outputfileExtension="jpg")
            return time.perf_counter() - t0
        run(os.path.join(root, "warm"))                                   # workspaces, textures, tables
        total = run(os.path.join(root, "out"))
        S._load_sequence, S._imwrite = orig_load, orig_write
        got = seen.get("offsets", [])
        ok = len(got) == len(offs) and all(abs(g[0] - t[0]) <= 1 and abs(g[1] - t[1]) <= 1 for g, t in zip(got, offs))
        results = sorted(os.listdir(os.path.join(root, "out")))
        res = cv2.imread(os.path.join(root, "out", results[0]), cv2.IMREAD_UNCHANGED) if results else None
        line = {
            "what": "%d x %d serpentine grid of %d^2 %s JPEG tiles (q92) -> offsets -> %s mosaic -> .jpg, through Stitcher.imageSetStitchWithMutiple"
                    % (a.rows, a.cols, a.size, "gray" if a.gray else "colour", a.fuse),
            "tiles": n, "tiles_per_s": n / total, "seconds": total,
            "split_s": {"decode": split["decode"], "align": total - sum(split.values()), "mosaic": split["mosaic"], "encode": split["encode"]},
            "offsets_within_1px": bool(ok), "pairs_found": len(got), "result_files": results,
            "result_shape": None if res is None else list(res.shape), "options": a.options,
        }
        if a.cpu_pairs > 0:
            from oracle import blend_oracle as bo
            from oracle import surf
            m = min(a.cpu_pairs + 1, n)
            files = sorted(os.listdir(d))[:m]
            t0 = time.perf_counter()
            gray = [cv2.imdecode(np.fromfile(os.path.join(d, f), np.uint8), 0) for f in files]
            gray = [cv2.imdecode(np.fromfile(os.path.join(d, f), np.uint8), 0) for f in files]      # the reference decodes twice (Stitcher.py:68-69)
            col = [cv2.imdecode(np.fromfile(os.path.join(d, f), np.uint8), 0 if a.gray else 1) for f in files]
            t_dec = time.perf_counter() - t0
            L = int(np.floor(a.size * 0.2))
            t0 = time.perf_counter()
            for k in range(m - 1):                                         # candidate i = 1 in the true direction only: a lower bound
                dr, dc = offs[k]
                if abs(dr) > abs(dc):
                    ra, rb = (gray[k][a.size - L:], gray[k + 1][:L]) if dr > 0 else (gray[k][:L], gray[k + 1][a.size - L:])
                else:
                    ra, rb = (gray[k][:, a.size - L:], gray[k + 1][:, :L]) if dc > 0 else (gray[k][:, :L], gray[k + 1][:, a.size - L:])
                ra, rb = np.ascontiguousarray(ra), np.ascontiguousarray(rb)
                mf = int(0.01 * ra.size)
                kA, dA = surf.detect_and_compute(ra, 100, 4, 3, False, False, 0)
                kB, dB = surf.detect_and_compute(rb, 100, 4, 3, False, False, 0)
                mt = cv2.BFMatcher().knnMatch(dA, dB, k=2)
                good = np.array([[p[0].trainIdx, p[0].queryIdx] for p in mt if len(p) == 2 and p[0].distance < 0.75 * p[1].distance], np.int32).reshape(-1, 2)
                surf.offset_by_mode(kA, kB, good, 3)
            t_al = time.perf_counter() - t0
            t0 = time.perf_counter()
            mosaic = bo.mosaic(np.stack(col), [list(o) for o in offs[:m - 1]], a.fuse)
            t_mo = time.perf_counter() - t0
            t0 = time.perf_counter()
            cv2.imwrite(os.path.join(root, "cpu.jpg"), mosaic)
            t_en = time.perf_counter() - t0
            per_tile = t_dec / m + t_al / max(m - 1, 1) + t_mo / m + t_en / m
            line["cpu_port"] = {"tiles_per_s": 1.0 / per_tile, "sample_tiles": m, "threads": surf.num_threads(), "kind": "port",
                                "split_s_per_tile": {"decode": t_dec / m, "align": t_al / max(m - 1, 1), "mosaic": t_mo / m, "encode": t_en / m},
                                "what": "cv2.imdecode x3 per tile, oracle C SURF (64-d, no cap: the cv2-CPU parameter set Main.py selects) + cv2 BFMatcher + "
                                        "vote port on the ROI of candidate i = 1 in the true direction, NumPy port of the paste / blend loop, cv2.imwrite"}
        print(json.dumps(line), flush=True)
    finally:
        if a.keep:
            print("kept", root, file=sys.stderr)
        else:
            shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
