"""Validated selection of kernel variants (include/vfsms.h VFSMS_OPT_*).

Every variant of an option computes the same result as the default schedule; which one is faster depends on the workload
shape.  `select()` runs a PROBE IN A SUBPROCESS on this rank's GPU: the default schedule and each variant execute the same
alignment step (device-resident tiles, the shape the caller is about to run) plus the single-image SURF entry point, and a
variant is eligible only if
  * keypoints (x, y, size, angle, response, octave, laplacian) and descriptors are BIT-IDENTICAL to the default's, and
  * the per-pair result table (status, offset, votes, keypoint and match counts, flags) is identical,
and it is selected only if it is also faster.  The chosen combination is validated again as a whole.  The subprocess
isolates the caller from anything a variant could do wrong (CUDA error, hang -> timeout): on any failure the answer is
"defaults".  Nothing here touches oracle/ or the CPU: the comparison is GPU default vs GPU variant.

    python -m imagestitch_b200.autotune --device 0 --pairs 32 --size 2048      # prints one JSON line
"""
import json
import os
import subprocess
import sys
import time

CANDIDATES = (("describe", 2), ("describe", 3), ("describe", 4), ("describe", 5), ("describe", 6), ("describe", 7), ("describe", 8),
              ("sort", 1), ("lpt", 1), ("lpt", 2), ("lpt", 3))
DEFAULTS = {"describe": 2, "sort": 1, "lpt": 2}
MIN_GAIN = 0.01          # a variant must be at least this much faster (fraction of the step) to be selected


def _probe(device, pairs, size, overlap, reps):
    import numpy as np
    import torch
    from . import gpu, synth
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    roi = int(np.floor(size * 0.2))
    A, B, _ = synth.pair_batch_torch(seed=777, n_pairs=pairs, size=size, overlap=overlap, device=dev)
    res = torch.zeros((pairs, 8), dtype=torch.int32, device=dev)
    singles = [np.ascontiguousarray(A[k, size - roi:, :].cpu().numpy()) for k in range(min(2, pairs))]
    singles.append(np.ascontiguousarray(B[0, :roi, : size // 2 + 3].cpu().numpy()))          # odd width: other pitch / borders
    params = gpu.surf_params()

    def run(opts):
        for name, v in DEFAULTS.items():
            gpu.set_option(name, opts.get(name, v), device=device)
        res.zero_()
        gpu.align_strips_dev(A, B, res, roi, params=params)
        torch.cuda.synchronize(dev)
        table = res.cpu().numpy().copy()
        feats = [gpu.surf_detect_and_describe(im, params=params, device=device) for im in singles]
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            gpu.align_strips_dev(A, B, res, roi, params=params)
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
        return table, feats, best * 1e3

    def same(x, y):
        if not np.array_equal(x[0], y[0]):
            return False
        return all(np.array_equal(ka, kb) and np.array_equal(da, db) for (ka, da), (kb, db) in zip(x[1], y[1]))

    run({})                                              # warm-up: workspaces, textures
    ref = run({})
    report = {"default_ms": ref[2], "pairs": pairs, "size": size, "status_ok": int(ref[0][:, 0].sum())}
    chosen, best_ms = {}, {}
    for name, v in CANDIDATES:
        out = run({name: v})
        ok = same(ref, out)
        report["%s=%d" % (name, v)] = {"identical": bool(ok), "ms": out[2]}
        if ok and out[2] < ref[2] * (1.0 - MIN_GAIN) and out[2] < best_ms.get(name, 1e30):
            chosen[name] = v                             # several values of one option: the fastest identical one
            best_ms[name] = out[2]
    if chosen:
        out = run(chosen)
        ok = same(ref, out)
        report["combined"] = {"options": dict(chosen), "identical": bool(ok), "ms": out[2]}
        if not ok or out[2] >= ref[2]:
            chosen = {}
    report["selected"] = chosen
    return report


def select(device=0, pairs=32, size=2048, overlap=205, reps=3, timeout=150):
    """-> (options to set, report).  Never raises: any failure of the probe means defaults."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "imagestitch_b200.autotune", "--device", str(device), "--pairs", str(pairs), "--size", str(size),
           "--overlap", str(overlap), "--reps", str(reps)]
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    env.pop("VFSMS_OPTS", None)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)                                 # the probe is a plain single-GPU process
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=root)
    except subprocess.TimeoutExpired:
        return {}, {"error": "probe timed out after %d s" % timeout}
    except OSError as e:
        return {}, {"error": "probe could not start: %s" % e}
    line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), None)
    if r.returncode != 0 or line is None:
        return {}, {"error": "probe exited %d: %s" % (r.returncode, (r.stderr or r.stdout)[-300:])}
    try:
        report = json.loads(line)
    except ValueError:
        return {}, {"error": "probe output not JSON"}
    sel = {k: int(v) for k, v in report.get("selected", {}).items() if k in DEFAULTS}
    return sel, report


def apply(device=0, **kw):
    """select() and set the chosen options on this process's context of `device`.  -> report."""
    from . import gpu
    chosen, report = select(device=device, **kw)
    for name, v in chosen.items():
        gpu.set_option(name, v, device=device)
    return report


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--overlap", type=int, default=205)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(_probe(a.device, a.pairs, a.size, a.overlap, a.reps)), flush=True)


if __name__ == "__main__":
    main()
