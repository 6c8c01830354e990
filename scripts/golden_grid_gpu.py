"""GPU check against the reference's own golden vector: the 89 pair offsets the author left at Stitcher.py:87 for
demoImages/dendriticCrystal/1.  The 90 JPEG tiles cannot travel with the repository; the builder copies them next to the repo for
one GPU call (directory given on the command line), this script decodes them with the library's JPEG decoder into the device tile
stack, runs the batched incremental search (sharding.evaluate_shard_batched, GPU-SURF parameters of ImageUtility.py:23-28 and --
second run -- the CPU parameter set, 64-d without keypoint cap) and writes the comparison to profiles/r02/.

    python scripts/golden_grid_gpu.py <dir with 1-001.jpg ... 1-090.jpg> [out.json]
"""
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from imagestitch_b200 import gpu, sharding


def main():
    src = sys.argv[1]
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02", "dendritic_89_pairs_gpu.json")
    files = sorted(glob.glob(os.path.join(src, "*.jpg")))
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "dendritic_offsets.json")))["golden_Stitcher_py_87"]
    assert len(files) == len(golden) + 1, (len(files), len(golden))
    datas = [open(f, "rb").read() for f in files]
    rows, cols = gpu.jpeg_info(datas[0])[:2]
    t0 = time.perf_counter()
    gpu.tiles_reserve(len(files), rows, cols)
    gpu.tiles_decode_jpeg(0, datas)
    t_decode = time.perf_counter() - t0
    roi_ratio = 0.2
    report = {"tiles": len(files), "shape": [rows, cols], "decode_s": t_decode, "decode_tiles_per_s": len(files) / t_decode, "runs": {}}
    for name, params in (("gpu_params_128d_ratio0.01", gpu.surf_params()),
                         ("cpu_params_64d_uncapped", gpu.surf_params(extended=False, keypoints_ratio=0.0))):
        evaluate = sharding.tiles_batch_evaluator(0, lambda i, d: int(i * roi_ratio * (rows if d in (1, 3) else cols)), params=params)
        t0 = time.perf_counter()
        table, calls = sharding.evaluate_shard_batched(evaluate, 0, len(golden), 1, 1, roi_ratio)
        results, requests, _ = sharding._missing_candidates(table, 0, 1, 1, roi_ratio)
        dt = time.perf_counter() - t0
        assert not requests
        offs, within1, exact = [], 0, 0
        for k, (st, i, d, off) in enumerate(results):
            o = sharding.roi_origin_back(off, (rows, cols), (rows, cols), i, d, roi_ratio) if st else None
            offs.append([bool(st), o, int(d)])
            if st:
                e = max(abs(o[0] - golden[k][0]), abs(o[1] - golden[k][1]))
                within1 += e <= 1; exact += e == 0
        report["runs"][name] = {"within_1px_of_golden": int(within1), "exact": int(exact), "pairs": len(golden), "seconds": dt,
                                "pairs_per_s": len(golden) / dt, "device_calls": calls, "offsets": offs}
        print(name, "within +-1 px of Stitcher.py:87:", within1, "/", len(golden), "exact", exact, "%.2f s" % dt, flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(report, open(out_path, "w"), indent=1)
    bad = [n for n, r in report["runs"].items() if r["within_1px_of_golden"] != len(golden)]
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
