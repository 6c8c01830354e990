#!/usr/bin/env python
"""bench.py -- pairwise alignments/s on 2048^2 tiles (BASELINE.json metric), one process per GPU.

A "step" = one pass of the hot path (2 x ROI SURF detect+describe -> kNN(2) match + ratio -> offset vote, i.e. one
body of Stitcher.calculateOffsetForFeatureSearchIncre succeeding at i=1 in the starting direction, roiRatio 0.2,
GPU-SURF parameters of ImageUtility.py:23-28) over a batch of synthetic tile pairs.
  value : whole-job pairs/s with tiles resident in HBM (device timing, max over ranks)
  e2e   : same metric through the public host API (gpu.align_batches: pinned host ROIs -> H2D -> kernels -> D2H, the copy of
          batch k+1 in flight while batch k runs; gpu.align_batch, one blocking call per batch, is timed beside it)
  --impl reference : the CPU port of the same path (oracle SURF on all host cores + cv2 BFMatcher + vote port)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun exports OMP_NUM_THREADS=1; the CPU arms (reference / cpu_baseline) must see every host core, and NCCL's
# version banner must not precede the single JSON line on stdout.
def _cgroup_cpu_quota():
    """CPUs' worth of CFS quota of this container (cgroup v2 cpu.max / v1 cfs_quota_us), or None when unlimited."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return float(q) / float(per)
    except (OSError, ValueError):
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return q / per
    except (OSError, ValueError):
        pass
    return None


def _physical_cores(allowed):
    """Distinct (socket, core) pairs among the allowed logical CPUs (hyper-thread siblings counted once)."""
    cores, cur = set(), {}
    try:
        for line in open("/proc/cpuinfo"):
            if ":" in line:
                k, v = [t.strip() for t in line.split(":", 1)]
                cur[k] = v
            elif cur:
                if int(cur.get("processor", -1)) in allowed:
                    cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                cur = {}
        if cur and int(cur.get("processor", -1)) in allowed:
            cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
    except (OSError, ValueError):
        return None
    return len(cores) or None


def host_cores():
    """Threads the CPU arms use: the physical cores this process may run on, capped by the container's CPU quota.
    (os.cpu_count() over-reports inside a cpuset; one busy-waiting OpenMP thread per hyper-thread or beyond the quota
    made the reference arm 12x slower than the same code with one thread per core.)"""
    try:
        allowed = set(os.sched_getaffinity(0))
    except AttributeError:
        allowed = set(range(os.cpu_count() or 1))
    n = len(allowed)
    phys = _physical_cores(allowed)
    if phys:
        n = min(n, phys)
    quota = _cgroup_cpu_quota()
    if quota:
        n = min(n, max(1, int(quota)))
    return max(1, n)


if "--impl" in sys.argv and "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")       # idle OpenMP workers must not spin while cv2's own pool runs the matcher
if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROI_RATIO = 0.2
TILE = 2048
OVERLAP = 205
METRIC = "pairwise_alignments_per_s_2048sq_surf"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_pipeline_pair(roiA, roiB, mf):
    """CPU port of one alignment: oracle SURF (OpenMP) + cv2 BFMatcher kNN(2) (the call the reference makes,
    ImageUtility.py:288-289) + ratio test + vote port."""
    import cv2
    from oracle import surf
    kA, dA = surf.detect_and_compute(roiA, 100.0, 4, 3, True, False, mf)
    kB, dB = surf.detect_and_compute(roiB, 100.0, 4, 3, True, False, mf)
    matcher = cv2.DescriptorMatcher_create("BruteForce")
    raw = matcher.knnMatch(dA, dB, 2)
    m = np.array([(r[0].trainIdx, r[0].queryIdx) for r in raw if len(r) == 2 and r[0].distance < r[1].distance * 0.75], np.int32).reshape(-1, 2)
    return surf.offset_by_mode(kA, kB, m, 3)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import cv2
    from imagestitch_b200 import synth
    from oracle import surf
    cv2.setNumThreads(host_cores())
    L = int(np.floor(TILE * ROI_RATIO))
    mf = int(0.01 * L * TILE)
    n_sample = 2
    tiles, offs = [], []
    for p in range(n_sample):
        A, B, off = synth.pair(seed=1234 + p, size=TILE, overlap=OVERLAP, direction=1)
        tiles.append((np.ascontiguousarray(A[TILE - L:]), np.ascontiguousarray(B[:L]))); offs.append(off)
    for _ in range(args.warmup):
        cpu_pipeline_pair(tiles[0][0], tiles[0][1], mf)
    t0 = time.perf_counter()
    ok = 0
    for _ in range(args.steps):
        for (a, b), off in zip(tiles, offs):
            st, o, v = cpu_pipeline_pair(a, b, mf)
            ok += int(st and abs(o[0] + TILE - L - off[0]) <= 1 and abs(o[1] - off[1]) <= 1)
    dt = time.perf_counter() - t0
    val = args.steps * n_sample / dt
    cores = surf.num_threads()
    # real-cv2 context numbers on the same ROI pair through the reference's own calls (ImageUtility.py:256-302:
    # SIFT_create().detectAndCompute + BFMatcher.knnMatch(k=2) + ratio 0.75; ORB_create(5000, 1.2, 8, 31, 0, 2, 0, 31, 20) + BFMatcher(
    # NORM_HAMMING).match) + the vote port -- and the probe for a contrib / non-free cv2 that could run the reference's SURF itself
    cv2_ctx = {"cv2_version": cv2.__version__}
    try:
        cv2.xfeatures2d.SURF_create()
        cv2_ctx["contrib_surf_available"] = True
    except Exception as e:                                                             # noqa: BLE001
        cv2_ctx["contrib_surf_available"] = False
        cv2_ctx["contrib_surf_probe"] = "%s: %s" % (type(e).__name__, str(e)[:120])
    a, b = tiles[0]

    def _vote(kA, kB, m):
        return surf.offset_by_mode(np.float32([k.pt for k in kA]).reshape(-1, 2), np.float32([k.pt for k in kB]).reshape(-1, 2), m, 3)
    try:
        t0 = time.perf_counter()
        sift = cv2.SIFT_create()
        kA, dA = sift.detectAndCompute(a, None); kB, dB = sift.detectAndCompute(b, None)
        raw = cv2.DescriptorMatcher_create("BruteForce").knnMatch(dA, dB, 2)
        m = np.array([(r[0].trainIdx, r[0].queryIdx) for r in raw if len(r) == 2 and r[0].distance < r[1].distance * 0.75], np.int32).reshape(-1, 2)
        st, o, v = _vote(kA, kB, m)
        cv2_ctx["sift_pairs_per_s"] = 1.0 / (time.perf_counter() - t0)
        cv2_ctx["sift_ok"] = bool(st and abs(o[0] + TILE - L - offs[0][0]) <= 1 and abs(o[1] - offs[0][1]) <= 1)
        t0 = time.perf_counter()
        orb = cv2.ORB_create(5000, 1.2, 8, 31, 0, 2, 0, 31, 20)
        kA, dA = orb.detectAndCompute(a, None); kB, dB = orb.detectAndCompute(b, None)
        mm = cv2.BFMatcher(cv2.NORM_HAMMING).match(dA, dB)
        m = np.array([(x.trainIdx, x.queryIdx) for x in sorted(mm, key=lambda x: x.queryIdx)], np.int32).reshape(-1, 2)
        st, o, v = _vote(kA, kB, m)
        cv2_ctx["orb_pairs_per_s"] = 1.0 / (time.perf_counter() - t0)
        cv2_ctx["orb_ok"] = bool(st and abs(o[0] + TILE - L - offs[0][0]) <= 1 and abs(o[1] - offs[0][1]) <= 1)
        cv2_ctx["what"] = "one ROI pair (409x2048) of this workload, cv2 on %d threads: SIFT / ORB detect+describe, BFMatcher, vote port" % cv2.getNumThreads()
    except Exception as e:                                                             # noqa: BLE001
        cv2_ctx["error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "synthetic 2048x2048 grayscale pair, SURF detect+describe+match+vote (ROI 409x2048, GPU-SURF params)",
                      "pairs_per_step": n_sample, "correct_pairs": ok, "cv2_threads": cv2.getNumThreads(), "host_cpus": host_cores()},
           "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port",
                            "sample": "%d pairs/step x %d steps: oracle C SURF (OpenMP %d thr) + cv2 BFMatcher knnMatch + ratio + vote port" % (n_sample, args.steps, cores),
                            "cv2_context": cv2_ctx},
           "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def aux_probe_main(path, device):
    """Child of run_aux_probe: the paths that have never run on hardware (device Huffman decoding, device JPEG encoder), each reported
    as a block or an {"error": ...}.  Prints one JSON line."""
    import cv2
    import torch
    from imagestitch_b200 import gpu
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    stream = torch.cuda.Stream(dev)
    tiles_h = np.load(path)
    n_t = tiles_h.shape[0]
    reps = 3
    out = {}
    try:
        files = [cv2.imencode(".jpg", t, [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes() for t in tiles_h]
        stack = torch.empty(tiles_h.shape, dtype=torch.uint8, device=dev)
        gpu.set_option("entropy", 0, device=device)
        gpu.jpeg_decode_gray_dev(files, stack, device=device, stream=stream)           # host entropy stage: the comparison
        gpu.set_option("entropy", 1, device=device)
        stack2 = torch.empty_like(stack)
        gpu.jpeg_decode_gray_dev(files, stack2, device=device, stream=stream)
        t0 = time.perf_counter()
        for _ in range(reps):
            gpu.jpeg_decode_gray_dev(files, stack2, device=device, stream=stream)      # synchronous on return
        dt_d = (time.perf_counter() - t0) / reps
        out["device_entropy"] = {"tiles_per_s": n_t / dt_d, "identical_to_host_stage": bool(torch.equal(stack, stack2)), "default": True,
                                 "sync_passes": gpu.jpeg_last_entropy_passes(device=device),
                                 "what": "same files, Huffman decoding on the device (self-synchronising 1024-bit subsequences): H2D of the unstuffed scan only"}
    except Exception as e:                                                             # noqa: BLE001
        out["device_entropy"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    print(json.dumps(out), flush=True)          # kept if the encoder below takes the process down
    try:
        t4 = torch.from_numpy(tiles_h[:4]).to(dev)                                     # four tiles -> one BGR canvas twice their side
        gray = torch.cat([torch.cat([t4[0], t4[1]], dim=1), torch.cat([t4[2], t4[3]], dim=1)], dim=0)
        canvas = torch.stack([gray, torch.roll(gray, 5, 0), 255 - torch.roll(gray, 9, 1)], dim=2).contiguous()
        torch.cuda.synchronize(dev)
        data = gpu.jpeg_encode_dev(canvas, device=device, stream=stream)               # warm-up: workspaces, tables
        t0 = time.perf_counter()
        for _ in range(reps):
            data = gpu.jpeg_encode_dev(canvas, device=device, stream=stream)           # synchronous on return
        dt_e = (time.perf_counter() - t0) / reps
        host_img = canvas.cpu().numpy()
        t0 = time.perf_counter()
        ref_bytes = cv2.imencode(".jpg", host_img)[1].tobytes()
        dt_c = time.perf_counter() - t0
        mp = host_img.shape[0] * host_img.shape[1] / 1e6
        out["output_encode"] = {"mpix_per_s": mp / dt_e, "cv2_imencode_mpix_per_s_1_thread": mp / dt_c, "identical_to_cv2": bool(data == ref_bytes),
                                "jpeg_bytes": len(data),
                                "what": "%dx%d BGR canvas in HBM -> baseline JPEG q95 4:2:0 (colour conversion, FDCT, quantisation, Huffman coding, byte "
                                        "stuffing on the device; D2H of the compressed stream only); wall clock" % host_img.shape[:2]}
    except Exception as e:                                                             # noqa: BLE001
        out["output_encode"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    print(json.dumps(out), flush=True)


def run_aux_probe(tiles_h, device, timeout=240):
    """aux_probe_main in a subprocess on `tiles_h` ([n, rows, cols] u8).  -> its last JSON line, or {"error": ...}; whatever happens
    in the child (CUDA error, crash, hang -> timeout) stays there."""
    import tempfile
    fd, path = tempfile.mkstemp(suffix=".npy")
    os.close(fd)
    out, err = "", None
    try:
        np.save(path, tiles_h)
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
            env.pop(k, None)
        cmd = [sys.executable, os.path.abspath(__file__), "--aux-probe", path, "--aux-device", str(device)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=os.path.dirname(os.path.abspath(__file__)))
            out = r.stdout
            if r.returncode != 0:
                err = "probe exited %d: %s" % (r.returncode, (r.stderr or "")[-300:])
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
            err = "probe timed out after %d s" % timeout
        except OSError as e:
            err = "probe could not start: %s" % e
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    res = {}
    for ln in reversed(out.splitlines()):
        if ln.startswith("{"):
            try:
                res = json.loads(ln)
                break
            except ValueError:
                continue
    if err:
        res["error"] = err
    return res


# ---------------------------------------------------------------- extra blocks of the bench line (BASELINE.json configs[2..4])
C4_GRID = (9, 10)          # 90 tiles, 89 consecutive pairs: the shape of demoImages/dendriticCrystal (configs[3])
C4_SEED = 4040


def _guard(fn):
    """An auxiliary block must never take the headline line down: its failure is reported inside the block."""
    def run(*a, **kw):
        try:
            return fn(*a, **kw)
        except Exception as e:                                                         # noqa: BLE001
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    return run


@_guard
def c4_block(rank, world, local, dev, steps):
    """configs[3], STRONG scaling: the 89 consecutive pairs of a 90-tile serpentine grid (synthetic at 2048^2: the demo JPEGs cannot
    ship), contiguous pair shards, tiles resident on their owner's GPU, candidates evaluated in batched rounds (vfsms_tiles_align),
    all_gather of the candidate tables + replay of the reference's search order (Stitcher.py:306-367) on every rank, misses sent back
    to their owners -- all inside the timed region."""
    import hashlib
    import torch
    import torch.distributed as dist
    from imagestitch_b200 import gpu, sharding, synth
    n_rows, n_cols = C4_GRID
    n_pairs = n_rows * n_cols - 1
    ranges = sharding.partition_pairs(n_pairs, world)
    s, e = ranges[rank]
    count = e - s + 1 if e > s else 0
    true_off = None
    if count:
        tiles, true_off = synth.sequence_torch(C4_SEED, n_rows, n_cols, TILE, OVERLAP, dev, first=s, count=count)
        gpu.tiles_reserve(count, TILE, TILE, device=local)
        gpu.tiles_upload(0, tiles.cpu().numpy(), device=local)            # outside the timed region: the stack is the HBM-resident input
        del tiles
        torch.cuda.empty_cache()
    else:
        _, true_off = synth.serpentine_origins(n_rows, n_cols, TILE, OVERLAP, C4_SEED)
    params = gpu.surf_params()
    evaluate = sharding.tiles_batch_evaluator(s, lambda i, d: int(i * ROI_RATIO * TILE), params=params, device=local)

    def run():
        return sharding.align_sequence_sharded_batched(evaluate, n_pairs, 1, 1, ROI_RATIO, rank, world, device=dev if world > 1 else None)
    run()                                                                              # warm-up: workspaces, plans
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        results, stats = run()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / steps
    offsets, ok = [], 0
    for k, (st, i, d, off) in enumerate(results):
        o = sharding.roi_origin_back(off, (TILE, TILE), (TILE, TILE), i, d, ROI_RATIO) if st else [0, 0]
        offsets.append([int(st), int(i), int(d), int(o[0]), int(o[1])])
        ok += int(bool(st) and abs(o[0] - int(true_off[k][0])) <= 1 and abs(o[1] - int(true_off[k][1])) <= 1)
    calls = torch.tensor([stats["device_calls"]], device=dev)
    if world > 1:
        dist.all_reduce(calls, op=dist.ReduceOp.MAX)
    return {"workload": "configs[3] shape: %d-tile serpentine grid (%dx%d tiles of %dx%d, synthetic), %d consecutive pairs, incremental SURF search "
                        "(roiRatio 0.2, direction 1, directIncre 1) sharded contiguously over %d GPU(s); gather + exact replay inside the timed region"
                        % (n_rows * n_cols, n_rows, n_cols, TILE, TILE, n_pairs, world),
            "scaling": "strong", "pairs": n_pairs, "pairs_per_s": n_pairs / sec, "ms_per_sequence": sec * 1e3, "runs_timed": steps,
            "within_1px_of_truth": "%d/%d" % (ok, n_pairs), "offsets_sha1": hashlib.sha1(json.dumps(offsets).encode()).hexdigest()[:16],
            "device_calls_max_rank": int(calls.item()), "replay_extra_rounds": stats["extra_rounds"], "replay_on_demand": stats["on_demand"],
            "collective": "all_gather of int32 candidate tables [pairs, 3, 4, 4] (NCCL)" if world > 1 else None}


@_guard
def phase_block(local, dev, stream, peaks):
    """configs[2]: cv2.phaseCorrelate(float64) replacement (Stitcher.py:230) on device-resident ROIs, CUDA events on the launching
    stream; the ROI of the incremental search on a 4096^2 pair (819 x 4096) and the full frame."""
    import cv2
    import torch
    from imagestitch_b200 import gpu, synth
    cv2.setNumThreads(host_cores())
    out = {}
    base = synth.canvas_torch(77, 4096 + 64, 4096 + 64, dev)
    a_full = base[10:10 + 4096, 20:20 + 4096].clamp(0, 255).to(torch.uint8).contiguous()
    b_full = base[15:15 + 4096, 8:8 + 4096].clamp(0, 255).to(torch.uint8).contiguous()          # b = a shifted by (+5, -12)
    res = torch.zeros(3, dtype=torch.float64, device=dev)
    for name, rows in (("roi_819x4096", 819), ("full_4096x4096", 4096)):
        a, b = a_full[:rows], b_full[:rows]
        M, N = cv2.getOptimalDFTSize(rows), cv2.getOptimalDFTSize(4096)
        with torch.cuda.stream(stream):
            for _ in range(3):
                gpu.phase_correlate_dev(a, b, res, stream=stream)
            reps = 20 if rows < 4096 else 8
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                gpu.phase_correlate_dev(a, b, res, stream=stream)
            e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        sx, sy, resp = (float(v) for v in res.cpu())
        ah, bh = a.cpu().numpy(), b.cpu().numpy()
        t0 = time.perf_counter()
        (cx, cy), cresp = cv2.phaseCorrelate(np.float64(ah), np.float64(bh))
        cv_ms = (time.perf_counter() - t0) * 1e3
        # B_pc = 2 h w (u8 in) + 56 M N for an fp32 pipeline; the library computes in float64 like the reference: float terms x 2
        alg = 2.0 * rows * 4096 + 112.0 * M * N
        out[name] = {"pairs_per_s": 1e3 / ms, "ms_per_pair": ms, "dft": [M, N], "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "shift": [sx, sy], "response": resp,
                     "cv2_phaseCorrelate_ms": cv_ms, "cv2_threads": cv2.getNumThreads(),
                     "max_abs_diff_vs_cv2": max(abs(sx - cx), abs(sy - cy), abs(resp - cresp))}
    out["what"] = "float64 like cv2.phaseCorrelate: pad + convert, cuFFT D2Z x 2, cross-power normalisation, Z2D, peak + 5x5 centroid; inputs resident in HBM"
    return out


@_guard
def mosaic_block(rank, world, local, dev, peaks):
    """configs[4] shape: full stitch of a serpentine sequence of 2048^2 tiles with fadeInAndFadeOut (Stitcher.py:369-486,
    ImageFusion.py:192-244), tiles resident in the HBM stack; N > 1: the sequence is cut into bands, one per GPU
    (sharding.mosaic_sharded: frontier patches travel rank to rank, rank 0 composes)."""
    import hashlib
    import torch
    import torch.distributed as dist
    from imagestitch_b200 import gpu, sharding, synth
    n_rows, n_cols = 4, 6
    n = n_rows * n_cols
    _, true_off = synth.serpentine_origins(n_rows, n_cols, TILE, OVERLAP, 2025)
    offs = [[0, 0]] + [[int(o[0]), int(o[1])] for o in true_off]
    origins, rois, shape = sharding.rectify_offsets(offs, [(TILE, TILE)] * n)
    out = {"tiles": n, "canvas": [int(shape[0]), int(shape[1])], "method": "fadeInAndFadeOut"}
    if world == 1:
        tiles, _ = synth.sequence_torch(2025, n_rows, n_cols, TILE, OVERLAP, dev)
        host_tiles = tiles.cpu().numpy()
        del tiles
        gpu.tiles_reserve(n, TILE, TILE, device=local)
        gpu.tiles_upload(0, host_tiles, device=local)
        gpu.tiles_mosaic(0, n, origins, rois, offs, "fadeInAndFadeOut", shape, device=local)          # warm-up
        gpu.profile_read(reset=True, device=local); gpu.profile_enable(True, device=local)
        t0 = time.perf_counter()
        canvas = gpu.tiles_mosaic(0, n, origins, rois, offs, "fadeInAndFadeOut", shape, device=local)
        dt = time.perf_counter() - t0
        gpu.profile_enable(False, device=local)
        st = gpu.profile_read(reset=True, device=local)
        blend_ms = st["blend"][0] if "blend" in st else None
        roi_px = float(sum(int(r[2] - r[0]) * int(r[3] - r[1]) for r in rois[1:]))
        alg = 2.0 * TILE * TILE * n + 4.0 * roi_px                      # per tile 2 H W (tile in, canvas out) + 4 r c
        out.update({"tiles_per_s": n / dt, "ms": dt * 1e3, "includes": "paste + blend on the device canvas and the D2H of the mosaic (%.0f MB)" % (canvas.size / 1e6),
                    "device_ms": blend_ms, "algorithmic_bytes": alg,
                    "achieved_gbs_device": alg / (blend_ms * 1e-3) / 1e9 if blend_ms else None,
                    "frac_of_hbm_peak_device": alg / (blend_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if blend_ms else None,
                    "mosaic_sha1": hashlib.sha1(canvas.tobytes()).hexdigest()[:16]})
        # CPU beside it: the reference's paste/blend loop restated in NumPy (oracle/blend_oracle.py) on the first tiles of the same sequence
        from oracle import blend_oracle
        k = 6
        t0 = time.perf_counter()
        ref = blend_oracle.mosaic(host_tiles[:k], np.asarray(true_off[:k - 1]), "fadeInAndFadeOut")
        dtc = time.perf_counter() - t0
        sub = gpu.mosaic(host_tiles[:k], *sharding.rectify_offsets(offs[:k], [(TILE, TILE)] * k)[:2], offs[:k], "fadeInAndFadeOut", ref.shape, device=local)
        out["cpu_port"] = {"tiles_per_s": k / dtc, "sample": "first %d tiles of the sequence, NumPy restatement of getStitchByOffset + fuseByFadeInAndFadeOut, 1 thread" % k,
                           "identical_to_device": bool(np.array_equal(ref, sub))}
    else:
        ranges = sharding.partition_pairs(n, world)
        s, e = ranges[rank]

        def load(a, b):
            t, _ = synth.sequence_torch(2025, n_rows, n_cols, TILE, OVERLAP, dev, first=a, count=b - a)
            return t.cpu().numpy()
        render = sharding.gpu_band_renderer(local)
        mine = load(s, e) if e > s else None                            # tile generation is not part of the stitch
        sharding.mosaic_sharded(render, lambda a, b: mine, true_off, (TILE, TILE), "fadeInAndFadeOut", rank, world, device=dev, gather=True)   # warm-up: NCCL point-to-point set-up
        dist.barrier()
        t0 = time.perf_counter()
        res = sharding.mosaic_sharded(render, lambda a, b: mine, true_off, (TILE, TILE), "fadeInAndFadeOut", rank, world, device=dev, gather=True)
        torch.cuda.synchronize(dev)
        dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out.update({"tiles_per_s": n / float(dt.item()), "ms": float(dt.item()) * 1e3, "bands": world,
                    "includes": "H2D of each rank's tiles, band render, frontier patches rank to rank, D2H + gather of the bands, composition on rank 0",
                    "mosaic_sha1": hashlib.sha1(res.tobytes()).hexdigest()[:16] if rank == 0 else None})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs", type=int, default=32, help="tile pairs per GPU per step")
    ap.add_argument("--batches", type=int, default=3, help="distinct input batches rotated through (L2 hygiene)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c4 / phase / mosaic blocks (profiling runs)")
    ap.add_argument("--matcher", default=None, choices=["tc", "tc_1sm", "simt", "tc_bf16x3", "tc_f16x2", "tc_f16x1"], help="override the descriptor matcher kernel")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="kernel-schedule switch (include/vfsms.h VFSMS_OPT_*, e.g. describe=0, lpt=1); identical results, A/B timing.  Without it the "
                         "bench times the library's default schedule, the one the oracle tests run on")
    ap.add_argument("--no-autotune", action="store_true", help=argparse.SUPPRESS)      # accepted for old command lines: there is no autotune
    ap.add_argument("--aux-probe", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--aux-device", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.aux_probe:
        return aux_probe_main(args.aux_probe, args.aux_device)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from imagestitch_b200 import gpu, synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, NB = args.pairs, args.batches
    L = int(np.floor(TILE * ROI_RATIO))
    params = gpu.surf_params()          # GPU-SURF defaults, ImageUtility.py:23-28
    if args.matcher:
        gpu.set_matcher(args.matcher, device=local)
    for item in args.opt:
        name, value = item.split("=")
        gpu.set_option(name, int(value), device=local)
    variants = {name: gpu.get_option(name, device=local) for name in gpu.OPTIONS}

    # ---- synthetic input, resident in HBM: NB distinct batches of P tile pairs (2 x NB x P x 4 MiB)
    batches = []
    for nb in range(NB):
        A, B, offs = synth.pair_batch_torch(seed=2025 + 97 * rank + nb, n_pairs=P, size=TILE, overlap=OVERLAP, device=dev)
        batches.append((A, B, offs))
    results = [torch.zeros((P, 8), dtype=torch.int32, device=dev) for _ in range(NB)]
    stream = torch.cuda.Stream(dev)          # explicit non-NULL stream: kernels and timing events share it
    torch.cuda.synchronize(dev)

    def step(i):
        A, B, _ = batches[i % NB]
        gpu.align_strips_dev(A, B, results[i % NB], L, params=params, stream=stream)

    torch.cuda.set_stream(stream)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    gpu.profile_read(reset=True, device=local)
    gpu.profile_enable(True, device=local)
    clocks = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks.start()
    l0 = gpu.launch_count(local)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    if world > 1:       # the only collective of the path: gather of the per-pair result table
        gathered = [torch.empty_like(results[0]) for _ in range(world)]
        dist.all_gather(gathered, results[(args.steps - 1) % NB])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = gpu.launch_count(local) - l0
    clk = clocks.stop()
    gpu.profile_enable(False, device=local)
    stages = gpu.profile_read(reset=True, device=local)
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())

    # ---- correctness of what was timed
    ok = 0; tot = 0; nkp = []; nmatch = []
    for nb in range(min(NB, args.steps)):
        r = results[nb].cpu().numpy(); offs = batches[nb][2]
        for p in range(P):
            tot += 1
            ok += int(r[p, 0] == 1 and abs(r[p, 1] + TILE - L - offs[p, 0]) <= 1 and abs(r[p, 2] - offs[p, 1]) <= 1 and r[p, 7] == 0)
            nkp.append((r[p, 4], r[p, 5])); nmatch.append(r[p, 6])
    nkp = np.array(nkp, np.float64); mean_na, mean_nb = nkp[:, 0].mean(), nkp[:, 1].mean()

    # ---- e2e: public host API with pinned host ROIs (H2D + kernels + D2H inside the timed region).  gpu.align_batches is
    # the streaming form of gpu.align_batch: the copy of batch k+1 is in flight while batch k's kernels run (two input slots);
    # every step's ROIs cross PCIe and every step's results come back inside the timed region.  Two distinct pinned
    # batches alternate so that no upload could be skipped.  The one-call-per-batch form is timed beside it (e2e_sync).
    A0, B0, offs0 = batches[0]
    A1, B1, offs1 = batches[1 % NB]
    hostA = torch.empty((2, P, L, TILE), dtype=torch.uint8).pin_memory(); hostB = torch.empty_like(hostA).pin_memory()
    hostA[0].copy_(A0[:, TILE - L:, :]); hostB[0].copy_(B0[:, :L, :])
    hostA[1].copy_(A1[:, TILE - L:, :]); hostB[1].copy_(B1[:, :L, :])
    nA, nB_ = hostA.numpy(), hostB.numpy()
    e2e_steps = max(4, args.steps // 2)
    for _ in gpu.align_batches(((nA[k & 1], nB_[k & 1]) for k in range(3)), params=params, device=local):
        pass
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    res_steps = list(gpu.align_batches(((nA[k & 1], nB_[k & 1]) for k in range(e2e_steps)), params=params, device=local))
    torch.cuda.synchronize(dev)
    dt_e2e = time.perf_counter() - t0
    te = torch.tensor([dt_e2e], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * P * e2e_steps / float(te.item())
    res_host = res_steps[0]
    gpu.align_batch(nA[1], nB_[1], params=params, device=local)            # its own staging buffers are allocated on first use
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(3):
        res_sync = gpu.align_batch(nA[k & 1], nB_[k & 1], params=params, device=local)
    torch.cuda.synchronize(dev)
    e2e_sync_val = world * P * 3 / (time.perf_counter() - t0)
    e2e_same = bool(all(np.array_equal(res_steps[k], res_steps[k & 1]) for k in range(e2e_steps)) and np.array_equal(res_sync, res_steps[0]))
    nA, nB_ = nA[0], nB_[0]                # the first batch's ROIs serve the probes and the CPU baseline below
    e2e_ok = int(sum(int(r["status"] == 1 and abs(r["d_row"] + TILE - L - offs0[p, 0]) <= 1) for p, r in enumerate(res_host)))

    # ---- the other named configurations (auxiliary blocks; every rank takes part in the sharded ones)
    c4 = mosaic = phase = None
    if not args.no_extra:
        c4 = c4_block(rank, world, local, dev, 3)
        mosaic = mosaic_block(rank, world, local, dev, load_peaks())
        if rank == 0 and world == 1:
            phase = phase_block(local, dev, stream, load_peaks())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage (CUDA-event time inside the timed region)
    peaks = load_peaks()
    # mean window area of the descriptor stage (win = floor(21 * 1.2 * size / 9)), from the keypoints of one ROI of this workload
    kp_probe, _ = gpu.surf_detect_and_describe(nA[0], params=params, device=local)
    win = np.floor(np.float32(21) * (kp_probe[:, 2] * np.float32(1.2) / np.float32(9.0))).astype(np.int64)
    mean_win2 = float((win * win).mean()) if len(win) else 0.0
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "traffic_ncu.json")
    if os.path.exists(tpath):
        traffic_db = json.load(open(tpath))
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages.items() if v[1] > 0}
    dom = max(stage_ms, key=stage_ms.get)
    D = 128
    px = L * TILE
    imgs = 2 * P
    alg = {
        # B_det ~ 88.7 B/px (u8 in + integral w/r + det/trace w + det r) over all 2P ROI images
        "integral": ("hbm", (1 + 4) * px * imgs),
        "hessian_nms": ("hbm", (4 + 2 * 4 * 5 * 1.328 + 4 * 5 * 1.328) * px * imgs),
        # B_kp = 113*16*4 (orientation Haar gathers) + win^2 (u8 window) + 4*D + 28 per keypoint
        "orient_describe": ("hbm", sum((113 * 16 * 4 + mean_win2 + 4 * D + 28) * n for n in (mean_na, mean_nb)) * P),
        "match_knn2": ("tensor", 2.0 * mean_na * mean_nb * D * P),
        "match_tc": ("tensor", 2.0 * mean_na * mean_nb * D * P),
    }
    roof = None
    if dom in alg:
        bound, work = alg[dom]
        t = stage_ms[dom] * 1e-3
        if bound == "hbm":
            ach = work / t / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": traffic_db.get(dom, {}).get("dram_bytes_per_launch_P%d" % P), "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": work, "mean_window_area_px": mean_win2,
                    "note": "instruction-issue bound, not HBM-bound: bit-exact CPU-order arithmetic costs ~30 warp instructions per 32 window samples plus the INTER_AREA fold (profiles/r02); DRAM traffic per launch is a fraction of the algorithmic gather bytes (texture / L2 hits)"}
        else:
            ach = work / t / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic_db.get(dom, {}).get("dram_bytes_per_launch_P%d" % P),
                    "peak_source": peaks["source"] + " (sustained)",
                    "algorithmic_flops_per_launch": work,
                    "matcher_gbs": (4 * D * (mean_na + mean_nb) + 16 * mean_na) * P / t / 1e9}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None}

    # the matcher's own roofline (tensor pipe), reported beside the dominant stage
    matcher = None
    if "match_tc" in stage_ms:
        tm = stage_ms["match_tc"] * 1e-3
        fl = 2.0 * mean_na * mean_nb * D * P
        matcher = {"stage_ms": stage_ms["match_tc"], "algorithmic_tflops": fl / tm / 1e12, "frac_of_bf16_sustained": fl / tm / 1e12 / peaks["bf16_tflops_sustained"],
                   "gbs": (4 * D * (mean_na + mean_nb) + 16 * mean_na) * P / tm / 1e9,
                   "kernel": args.matcher or "tc", "exact_rescans_last_step": gpu.last_match_fallbacks(local), "queries_per_step": int(mean_na * P),
                   "error_bound_violations_last_step": gpu.last_match_bound_violations(local),
                   "executed_over_algorithmic_flops": {"tc_bf16x3": (3 * D + 16) / D, "tc_f16x2": (2 * D + 16) / D}.get(args.matcher, (D + 16) / D) if args.matcher not in ("simt",) else 1.0,
                   "note": "stage = operand rows + tcgen05 GEMM with top-4 epilogue + exact fp32 rescoring + exact rescan of flagged queries; "
                           "default operands: fp16 rows of K' = D + 16 columns (the norm columns ride in a 16-column k-step)"}
    surf_kps = (mean_na + mean_nb) * P * world / (sum(stage_ms.get(k, 0) for k in ("integral", "hessian_nms", "rank_sort", "validate_compact", "orient_describe")) * 1e-3)

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import surf
        mf = int(0.01 * L * TILE)
        sample = 3
        cpu_pipeline_pair(nA[0], nB_[0], mf)
        t0 = time.perf_counter()
        for p in range(sample):
            cpu_pipeline_pair(nA[p], nB_[p], mf)
        dtc = time.perf_counter() - t0
        cpu = {"value": sample / dtc, "unit": "pairs/s", "cores": surf.num_threads(), "kind": "port",
               "sample": "%d ROI pairs (409x2048) of this workload: oracle C SURF (OpenMP) + cv2 BFMatcher knnMatch + ratio + vote port" % sample}

    # ---- tile ingest: JPEG files in host memory -> u8 tiles resident in HBM, beside cv2.imdecode
    ingest = None
    encode = None
    if world == 1 and not args.no_cpu_baseline:
        import cv2
        n_t = min(16, P)
        tiles_h = batches[0][0][:n_t].cpu().numpy()
        files = [cv2.imencode(".jpg", t, [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes() for t in tiles_h]
        stack = torch.empty((n_t, TILE, TILE), dtype=torch.uint8, device=dev)
        entropy_default = gpu.get_option("entropy", device=local)
        gpu.set_option("entropy", 0, device=local)                                     # this block: the HOST entropy stage (device stage below)
        gpu.jpeg_decode_gray_dev(files, stack, device=local, stream=stream)            # warm-up: buffers, threads
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            gpu.jpeg_decode_gray_dev(files, stack, device=local, stream=stream)        # synchronous on return
        dt_b = (time.perf_counter() - t0) / reps
        gpu.set_option("entropy", entropy_default, device=local)
        t0 = time.perf_counter()
        ref = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_GRAYSCALE) for f in files[:8]]
        dt_c = (time.perf_counter() - t0) / 8
        exact = all(np.array_equal(stack[k].cpu().numpy(), ref[k]) for k in range(8))
        ingest = {"tiles_per_s": n_t / dt_b, "mpix_per_s": n_t * TILE * TILE / dt_b / 1e6, "bit_exact_vs_cv2": bool(exact),
                  "jpeg_bytes_per_tile": int(np.mean([len(f) for f in files])), "host_threads": min(len(os.sched_getaffinity(0)), int(_cgroup_cpu_quota() or 1 << 30), 32, n_t),
                  "cv2_imdecode_tiles_per_s_1_thread": 1.0 / dt_c,
                  "what": "%d synthetic 2048x2048 JPEG tiles (q92, single component) from host bytes to HBM-resident u8 tiles with option entropy=0: host Huffman threads + H2D of int16 coefficients + IDCT kernel; wall clock.  The library default is entropy=1: see device_entropy" % n_t}
        # option "entropy" = 1 (Huffman decoding on the device) and the device JPEG encoder were written without GPU access and verified
        # on the CPU emulation only: their first hardware run happens in a SUBPROCESS (own CUDA context, timeout), so neither a CUDA
        # error nor a hang there can take the headline line down
        aux = run_aux_probe(tiles_h, local)
        ingest["device_entropy"] = aux.get("device_entropy", {"error": aux.get("error", "no result")})
        encode = aux.get("output_encode", {"error": aux.get("error", "no result")})

    value = world * P * args.steps / (ms_max * 1e-3)
    out = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "synthetic 2048x2048 grayscale pair, SURF detect+describe+match+vote (configs[1]); ROI 409x2048 read in place from HBM-resident tiles",
                      "pairs_per_gpu_per_step": P, "global_pairs_per_step": P * world, "roi": [L, TILE], "surf": "thr100 oct4 layers3 128-d ratio0.01",
                      "l2": "%d distinct input batches rotated; per-step intermediate traffic > L2" % NB,
                      "mean_keypoints": [mean_na, mean_nb], "mean_matches": float(np.mean(nmatch)), "correct_pairs": "%d/%d" % (ok, tot),
                      "e2e_correct_pairs": "%d/%d" % (e2e_ok, P), "kernel_variants": variants},
           "clocks": clk, "gpu_launches": int(launches),
           "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * P * L * TILE), "d2h_bytes_per_step": int(P * 32),
                   "steps": e2e_steps, "api": "gpu.align_batches (vfsms_align_batch_upload / _run, two input slots)",
                   "one_call_per_batch_pairs_per_s": e2e_sync_val, "streamed_results_equal_one_call_results": e2e_same},
           "roofline": roof, "stages_ms_per_step": stage_ms, "matcher": matcher, "surf_keypoints_per_s": surf_kps}
    if c4:
        out["c4"] = c4
    if phase:
        out["phase"] = phase
    if mosaic:
        out["mosaic"] = mosaic
    if cpu:
        out["cpu_baseline"] = cpu
    if ingest:
        out["tile_ingest"] = ingest
    if encode:
        out["output_encode"] = encode
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
