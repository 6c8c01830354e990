"""Secondary measurement (not the driver's bench line): mosaic assembly + fade blend of a serpentine tile grid -- the shape of
BASELINE.json configs[4] (synthetic 2048^2 sequence, full stitch + fadeInAndFadeOut), scaled by --rows / --cols.

    python scripts/bench_mosaic.py --rows 6 --cols 8                       # one GPU: vfsms_mosaic_host / vfsms_tiles_mosaic
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 scripts/bench_mosaic.py --rows 6 --cols 8
                                                                           # bands over 4 GPUs (sharding.mosaic_sharded, NCCL p2p)
Prints one JSON line: tiles/s (host tiles in, host mosaic out), the blend's share against the HBM roofline
(2 H W + 4 r c bytes per gray tile), and the NumPy oracle of the reference's loop timed on a bounded sample
of the same grid (cpu_baseline, kind "port": oracle/blend_oracle.py restates Stitcher.getStitchByOffset + ImageFusion).
Every run checks the device mosaic of the sample against the oracle byte for byte.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_grid(seed, n_rows, n_cols, size, overlap):
    from imagestitch_b200 import synth
    origins, offsets = synth.serpentine_origins(n_rows, n_cols, size, overlap, seed)
    base = synth.canvas(seed, size + 64, size + 64)
    rng = np.random.default_rng(seed)
    tiles = np.empty((len(origins), size, size), np.uint8)
    for k in range(len(origins)):                       # content only matters through which pixels are > 0; cheap variety
        dy, dx = (int(v) for v in rng.integers(0, 64, 2))
        tiles[k] = base[dy:dy + size, dx:dx + size]
    return tiles, [[int(o[0]), int(o[1])] for o in offsets]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--cols", type=int, default=6)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--overlap", type=int, default=205)
    ap.add_argument("--method", default="fadeInAndFadeOut")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-tiles", type=int, default=6, help="tiles of the grid's first rows the CPU oracle is timed / checked on")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    import torch
    import torch.distributed as dist
    from imagestitch_b200 import gpu, sharding
    from oracle import blend_oracle as bo
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tiles, offs = make_grid(2025, args.rows, args.cols, args.size, args.overlap)
    n = len(tiles)
    origins, rois, shape = sharding.rectify_offsets([[0, 0]] + offs, [(args.size, args.size)] * n)
    pair = np.asarray([[0, 0]] + offs, np.int32)

    def run_once():
        if world == 1:
            return gpu.mosaic(tiles, origins, rois, pair, args.method, shape, device=local)
        return sharding.mosaic_sharded(sharding.gpu_band_renderer(local), lambda s, e: tiles[s:e], offs, (args.size, args.size),
                                       args.method, rank, world, device=dev)
    out = run_once()                                    # warm-up: canvas / scratch allocation
    if world > 1:
        dist.barrier()
    gpu.profile_read(reset=True, device=local); gpu.profile_enable(True, device=local)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = run_once()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / args.reps
    gpu.profile_enable(False, device=local)
    stages = gpu.profile_read(reset=True, device=local)
    tt = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    if rank != 0:
        dist.destroy_process_group()
        return

    # parity + CPU baseline on a bounded sample: the first cpu_tiles tiles of the same grid
    m = min(args.cpu_tiles, n)
    t0 = time.perf_counter()
    ref = bo.mosaic(tiles[:m], offs[:m - 1], args.method)
    cpu_dt = time.perf_counter() - t0
    o2, r2, s2 = sharding.rectify_offsets([[0, 0]] + offs[:m - 1], [(args.size, args.size)] * m)
    dev_small = gpu.mosaic(tiles[:m], o2, r2, pair[:m], args.method, s2, device=local)
    exact = bool(np.array_equal(dev_small, ref))

    blend_ms = stages.get("blend", (0.0, 0))[0] / max(args.reps, 1)
    roi_px = int(sum((int(r[2]) - int(r[0])) * (int(r[3]) - int(r[1])) for r in rois[1:]))
    alg_bytes = 2 * n * args.size * args.size + 4 * roi_px
    peak = 6650.0
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peak = float(json.load(open(ppath)).get("hbm_gbs", peak))
    line = {"metric": "mosaic_tiles_per_s_2048sq_fade", "value": n / dt, "unit": "tiles/s", "n_gpus": world, "reps": args.reps,
            "ms_per_mosaic": dt * 1e3, "data": "synthetic",
            "config": {"workload": "serpentine %d x %d grid of %d^2 gray tiles, overlap %d, %s; host tiles in, host mosaic out (BASELINE configs[4] shape)"
                                   % (args.rows, args.cols, args.size, args.overlap, args.method),
                       "canvas": [int(shape[0]), int(shape[1])], "roi_pixels_blended": roi_px,
                       "path": "vfsms_mosaic_host" if world == 1 else "sharding.mosaic_sharded over %d bands (vfsms_mosaic_band_host + NCCL p2p)" % world},
            "identical_to_oracle_on_sample": exact,
            "roofline": {"kernel": "blend (stats + plan + apply)", "bound": "hbm", "achieved": alg_bytes / (blend_ms * 1e-3) / 1e9 if blend_ms > 0 else None,
                         "peak": peak, "unit": "GB/s", "frac": (alg_bytes / (blend_ms * 1e-3) / 1e9 / peak) if blend_ms > 0 else None,
                         "traffic": None, "algorithmic_bytes_per_mosaic": alg_bytes, "blend_stage_ms": blend_ms,
                         "note": "blend stage time = CUDA events of this rank; the rest of the wall time is H2D of the tiles, paste copies and D2H of the canvas"},
            "cpu_baseline": {"value": m / cpu_dt, "unit": "tiles/s", "cores": 1, "kind": "port",
                             "sample": "first %d tiles of the same grid through oracle/blend_oracle.py (NumPy restatement of getStitchByOffset + ImageFusion)" % m}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
