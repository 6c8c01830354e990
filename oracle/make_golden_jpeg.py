"""Golden fixtures for the JPEG tile decode.  Run in the build container (needs /root/reference):

    python oracle/make_golden_jpeg.py

Copies two of the reference's demo tiles (data, not code) -- one single-component file and one YCbCr 4:2:0 file, as the
microscope software wrote them -- into tests/golden/ and records shape + SHA-256 of `cv2.imdecode(data, 0)`, the decode
the reference performs at Stitcher.py:68-69.  It also checks every demo JPEG of the reference against the oracle
restatement (oracle/jpeg_oracle.py) fed by the library's host entropy stage, and prints the tally quoted in DESIGN.md.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jpeg_oracle as jo          # noqa: E402
from imagestitch_b200 import gpu              # noqa: E402  (host-only entry points: no GPU needed)

REF = "/root/reference/demoImages"
PICK = {"jpeg_gray_zirconBSE.jpg": "zirconBSE/1/LF0117-16-17-LYV-361-16-17-LYV-361-04.jpg",
        "jpeg_ycc420_zirconREM.jpg": "zirconREM/1/17DQ56-1-11.jpg"}


def bgr_through_host_stage(data):
    """Colour decode = the library's host entropy stage (all components) + the oracle's IDCT / upsampling / colour conversion."""
    info = jo.parse(data)
    rows, cols, n = info["rows"], info["cols"], len(info["comps"])
    planes, samp = [], []
    for c in range(n):
        coef, quant, hv = gpu.jpeg_component_coefficients(data, c)
        planes.append(jo.idct_islow(coef, quant.astype(np.int32))); samp.append(hv)
    if n == 1:
        g = planes[0][:rows, :cols]
        return np.stack([g, g, g], -1)
    hmax, vmax = max(s[0] for s in samp), max(s[1] for s in samp)
    up = [jo.upsample(pl, hmax // h, vmax // v, -(-rows * v // vmax), -(-cols * h // hmax))[:rows, :cols] for pl, (h, v) in zip(planes, samp)]
    return jo.ycc_to_bgr(*up)


def main():
    out = {}
    for name, rel in PICK.items():
        dst = os.path.join(ROOT, "tests", "golden", name)
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
        data = np.fromfile(dst, np.uint8)
        img = cv2.imdecode(data, cv2.IMREAD_GRAYSCALE)
        bgr = cv2.imdecode(data, cv2.IMREAD_COLOR)
        out[name] = {"source": "demoImages/" + rel, "rows": int(img.shape[0]), "cols": int(img.shape[1]),
                     "sha256_of_cv2_imdecode_gray": hashlib.sha256(img.tobytes()).hexdigest(),
                     "sha256_of_cv2_imdecode_color": hashlib.sha256(bgr.tobytes()).hexdigest(), "cv2": cv2.__version__}
    files = sorted(glob.glob(REF + "/*/*/*.[jJ][pP][gG]"))
    bad = bad_c = 0
    for f in files:
        data = np.fromfile(f, np.uint8)
        ref = cv2.imdecode(data, cv2.IMREAD_GRAYSCALE)
        coef, quant = gpu.jpeg_luma_coefficients(data.tobytes())
        mine = jo.idct_islow(coef, quant.astype(np.int32))[:ref.shape[0], :ref.shape[1]]
        bad += not np.array_equal(ref, mine)
        bad_c += not np.array_equal(cv2.imdecode(data, cv2.IMREAD_COLOR), bgr_through_host_stage(data.tobytes()))
    out["_demo_sweep"] = {"files": len(files), "bit_exact_vs_cv2": len(files) - bad, "color_bit_exact_vs_cv2": len(files) - bad_c}
    with open(os.path.join(ROOT, "tests", "golden", "jpeg_cases.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
