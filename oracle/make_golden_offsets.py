"""Generates tests/golden/dendritic_offsets.json and the real-image ROI fixtures -- run in the build container only.

Drives the UNMODIFIED reference Stitcher (via oracle/reference_shims.py) with cv2.xfeatures2d.SURF_create bound to
the C restatement (oracle/surf_oracle.c) over demoImages/dendriticCrystal/1 (90 tiles) and records, per pair, the
offset it returns next to the author's golden list (Stitcher.py:87).  Also stores lossless crops of the ROI strips
of a few pairs so that GPU-box tests have real micrograph input without /root/reference.
"""
import glob
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shims as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    S, U, F = rs.import_reference()
    golden = rs.golden_offsets()
    files = sorted(glob.glob(os.path.join(rs.REFERENCE, "demoImages/dendriticCrystal/1/*.jpg")))
    assert len(files) == 90 and len(golden) == 89
    St = S.Stitcher
    St.featureMethod = "surf"; St.isGPUAvailable = False; St.searchRatio = 0.75; St.offsetCaculate = "mode"
    St.offsetEvaluate = 3; St.roiRatio = 0.2; St.direction = 1; St.directIncre = 1
    st = St()
    res = []
    t0 = time.time()
    imgs = [cv2.imdecode(np.fromfile(f, dtype=np.uint8), cv2.IMREAD_GRAYSCALE) for f in files]
    dirs = []
    for k in range(89):
        dirs.append(int(St.direction if not hasattr(st, "direction") else st.direction))
        status, off = st.calculateOffsetForFeatureSearchIncre([imgs[k], imgs[k + 1]])
        res.append([bool(status), [int(off[0]), int(off[1])] if status else None, int(st.direction)])
        print(k, status, off, golden[k], flush=True)
    el = time.time() - t0
    within1 = sum(1 for r, g in zip(res, golden) if r[0] and abs(r[1][0] - g[0]) <= 1 and abs(r[1][1] - g[1]) <= 1)
    exact = sum(1 for r, g in zip(res, golden) if r[0] and r[1] == list(g))
    json.dump({"source": "reference Stitcher.calculateOffsetForFeatureSearchIncre + oracle SURF (64-d, CPU params)",
               "golden_Stitcher_py_87": golden, "oracle_surf_offsets": res, "within1": within1, "exact": exact,
               "seconds": el, "shape": list(imgs[0].shape)},
              open(os.path.join(OUT, "dendritic_offsets.json"), "w"))
    print("within1", within1, "exact", exact, "of 89; seconds", el)
    # ROI fixtures (direction 1 = bottom strip of A / top strip of B; direction 2 = right/left strips)
    h, w = imgs[0].shape
    L = int(np.floor(h * 0.2))
    for k in (0, 43):
        cv2.imwrite(os.path.join(OUT, "dendritic_%02d_A_dir1.png" % k), imgs[k][h - L:, :])
        cv2.imwrite(os.path.join(OUT, "dendritic_%02d_B_dir1.png" % k), imgs[k + 1][:L, :])
    iron = sorted(glob.glob(os.path.join(rs.REFERENCE, "demoImages/iron/1/*.jpg")))
    A = cv2.imdecode(np.fromfile(iron[0], dtype=np.uint8), 0); B = cv2.imdecode(np.fromfile(iron[1], dtype=np.uint8), 0)
    L = int(np.floor(A.shape[0] * 0.2))
    cv2.imwrite(os.path.join(OUT, "iron_A_dir1.png"), A[A.shape[0] - L:, :])
    cv2.imwrite(os.path.join(OUT, "iron_B_dir1.png"), B[:L, :])


if __name__ == "__main__":
    main()
