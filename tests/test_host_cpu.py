"""Host-side logic and the C-ABI surface -- no GPU needed."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from imagestitch_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.SO_PATH)
    header = open(os.path.join(ROOT, "include", "vfsms.h")).read()
    declared = set(re.findall(r"\b(vfsms_[a-z0-9_]+)\s*\(", header))
    declared -= {"vfsms_ctx"}
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(L, name), "libvfsms.so does not export %s" % name
    assert set(_lib.EXPORTS) <= declared


def test_no_cpu_fallback_without_device():
    from imagestitch_b200 import _lib
    L = _lib.load()
    if L.vfsms_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = ctypes.c_void_p()
    assert L.vfsms_create(0, ctypes.byref(h)) == -1 and b"no CPU fallback" in L.vfsms_last_error()
    from imagestitch_b200 import gpu
    with pytest.raises(gpu.VfsmsError):
        gpu.surf_detect_and_describe(np.zeros((32, 32), np.uint8))
    from imagestitch_b200.Stitcher import Stitcher
    with pytest.raises(gpu.VfsmsError):
        Stitcher().calculateOffsetForFeatureSearchIncre([np.zeros((64, 64), np.uint8)] * 2)


def test_plugin_package_imports_without_gpu_and_keeps_call_surface():
    from myGpuFeatures import myGpuFeatures as plugin
    import inspect
    assert list(inspect.signature(plugin.detectAndDescribeBySurf).parameters) == ["image", "hessianThreshold", "nOctaves", "nOctaveLayers", "isExtended", "keypointsRatio", "isUpright"]
    assert len(inspect.signature(plugin.detectAndDescribeByOrb).parameters) == 11
    assert list(inspect.signature(plugin.matchDescriptors).parameters) == ["descA", "descB", "featureType", "param"]
    out = plugin._pack(np.array([[1.5, 2.5, 0, 0, 0, 0, 0, 0]], np.float32), np.arange(128, dtype=np.float32)[None])
    assert out.shape == (1, 128, 2) and out[0, 0, 0] == 1.5 and out[0, 1, 0] == 2.5 and out[0, 5, 1] == 5 and out[0, 2, 0] == 0


def test_reference_call_surface_names():
    """Main.py sets these class attributes and calls these methods (Main.py:5-20)."""
    from Stitcher import Stitcher
    import ImageFusion
    import ImageUtility
    for attr in ("featureMethod", "isColorMode", "isGPUAvailable", "isEnhance", "isClahe", "searchRatio", "offsetCaculate", "offsetEvaluate",
                 "roiRatio", "fuseMethod", "direction", "directIncre", "outputAddress", "phaseResponseThreshold", "tempImageFeature",
                 "surfHessianThreshold", "surfNOctaves", "surfNOctaveLayers", "surfIsExtended", "surfKeypointsRatio", "surfIsUpright",
                 "orbNfeatures", "orbMaxDistance", "clipLimit", "tileSize"):
        assert hasattr(Stitcher, attr), attr
    for meth in ("imageSetStitchWithMutiple", "imageSetStitch", "flowStitch", "flowStitchWithMutiple", "calculateOffsetForFeatureSearch",
                 "calculateOffsetForFeatureSearchIncre", "calculateOffsetForPhaseCorrleateIncre", "calculateOffsetForPhaseCorrleate",
                 "getStitchByOffset", "fuseImage", "directionIncrease", "getROIRegionForIncreMethod", "detectAndDescribe", "matchDescriptors",
                 "getOffsetByMode", "getOffsetByRansac", "npToKpsAndDescriptors", "npToListForMatches", "npToListForKeypoints"):
        assert callable(getattr(Stitcher, meth)), meth
    assert issubclass(Stitcher, ImageUtility.Method) and issubclass(ImageFusion.ImageFusion, ImageUtility.Method)
    s = Stitcher()
    assert s.calculateOffsetForFeatureSearchIncre == s.calculateOffsetForFeatureSearchIncre     # bound-method comparison (Stitcher.py:70)


def test_direction_and_roi_logic():
    from imagestitch_b200.Stitcher import Stitcher
    from oracle import numpy_oracle as no
    s = Stitcher()
    Stitcher.directIncre = 1
    assert [s.directionIncrease(d) for d in (1, 2, 3, 4)] == [2, 3, 4, 1]
    Stitcher.directIncre = -1
    assert [s.directionIncrease(d) for d in (1, 2, 3, 4)] == [4, 1, 2, 3]
    Stitcher.directIncre = 0
    assert [s.directionIncrease(d) for d in (1, 4)] == [1, 4]          # Q5: only the initial direction is ever tried
    Stitcher.directIncre = 1
    img = np.arange(50 * 80, dtype=np.uint8).reshape(50, 80)
    for d in (1, 2, 3, 4):
        for order in ("first", "second"):
            for ratio in (0.2, 0.4, 0.6000000000000001):
                assert np.array_equal(s.getROIRegionForIncreMethod(img, d, order, ratio), no.roi_for_incre(img, d, order, ratio))
    assert s.getROIRegionForIncreMethod(img, 1, "first", 0.2).shape == (10, 80)
    assert not s.getROIRegionForIncreMethod(img, 2, "first", 0.2).flags["C_CONTIGUOUS"]


def test_incre_search_order_matches_reference_loop():
    """_incre_search visits candidates in the reference order and adds the ROI origin back (Stitcher.py:316-361)."""
    from imagestitch_b200.Stitcher import Stitcher
    s = Stitcher(); Stitcher.isPrintLog = False
    Stitcher.roiRatio = 0.2; Stitcher.direction = 1; Stitcher.directIncre = 1
    seen = []

    def ev(i, d):
        seen.append((i, d))
        return (i == 2 and d == 3, [5, -2])
    a = np.zeros((100, 200), np.uint8)
    st, off = s._incre_search([a, a], ev)
    assert seen == [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3)]
    assert st and off == [5 - (100 - int(2 * 0.2 * 100)), -2] and s.direction == 3
    Stitcher.direction = 1; Stitcher.isPrintLog = True
    if "direction" in s.__dict__:
        del s.__dict__["direction"]


def test_sharding_partition_and_replay_equals_sequential():
    from imagestitch_b200 import sharding as sh
    assert sh.partition_pairs(89, 8) == [(0, 12), (12, 23), (23, 34), (34, 45), (45, 56), (56, 67), (67, 78), (78, 89)]
    assert sh.max_i(0.2) - 1 == 3
    rng = np.random.default_rng(0)
    true_dir = np.array([1] * 14 + [2] + [3] * 14 + [2] + [1] * 10)           # serpentine
    n = len(true_dir)

    def evaluate(pair, i, d):
        ok = (d == true_dir[pair]) and (i >= 1 + (pair % 7 == 3))              # some pairs need the second ROI size
        spurious = (pair % 11 == 5 and d == 4 and i == 1)                      # a wrong direction that also "succeeds"
        return int(ok or spurious), 100 + pair, -pair, 9
    seq_table = sh.evaluate_shard(evaluate, 0, n, 1, 1, 0.2)
    seq, _ = sh.replay(seq_table.copy(), evaluate, 1, 1, 0.2)
    for world in (2, 3, 8):
        ranges = sh.partition_pairs(n, world)
        parts = [sh.evaluate_shard(evaluate, s, e, 1, 1, 0.2) for s, e in ranges]
        full = np.concatenate(parts, 0)
        out, on_demand = sh.replay(full, evaluate, 1, 1, 0.2)
        assert out == seq
        assert on_demand <= 4 * world


GLOO_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["VFSMS_ROOT"])
import torch.distributed as dist
from imagestitch_b200 import sharding as sh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
true_dir = [1] * 6 + [2] + [3] * 6 + [2] + [1] * 5
def evaluate(pair, i, d):
    return int(d == true_dir[pair]), 50 + pair, pair - 3, 7
out, on_demand = sh.align_sequence_sharded(evaluate, len(true_dir), 1, 1, 0.2, rank, world)
seq, _ = sh.replay(sh.evaluate_shard(evaluate, 0, len(true_dir), 1, 1, 0.2), evaluate, 1, 1, 0.2)
assert out == seq, (rank, out, seq)
if rank == 0:
    print("GLOO_OK", len(out), on_demand)
dist.destroy_process_group()
'''


def test_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, VFSMS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK 19" in r.stdout


def test_sharding_batched_rounds_equal_sequential():
    """evaluate_shard_batched fills, in rounds of grouped requests, a table on which the sequential walk needs nothing else."""
    from imagestitch_b200 import sharding as sh
    true_dir = np.array([1] * 9 + [2] + [3] * 9 + [2] + [1] * 9 + [2] + [3] * 4)
    n = len(true_dir)

    def evaluate(pair, i, d):
        ok = (d == true_dir[pair]) and (i >= 1 + (pair % 7 == 3))
        spurious = (pair % 11 == 5 and d == 4 and i == 1)
        return int(ok or spurious), 100 + pair, -pair, 9
    calls = []

    def batch_evaluate(pairs, i, dirs):
        assert len({d in (2, 4) for d in dirs}) == 1         # one strip shape per device call
        calls.append((tuple(zip(pairs, dirs)), i))
        return [evaluate(p, i, d) for p, d in zip(pairs, dirs)]
    for incre in (1, -1, 0):
        seq, _ = sh.replay(sh.evaluate_shard(evaluate, 0, n, 1, incre, 0.2), evaluate, 1, incre, 0.2)
        del calls[:]
        table, n_calls = sh.evaluate_shard_batched(batch_evaluate, 0, n, 1, incre, 0.2)
        out, requests, _ = sh._missing_candidates(table, 0, 1, incre, 0.2)
        assert not requests and out == seq
        assert n_calls == len(calls) <= 24                 # a handful of batched device calls instead of ~n single ones
        assert all(len(set(c[0])) == len(c[0]) for c in calls)
        on_path = int((sh.evaluate_shard(evaluate, 0, n, 1, incre, 0.2)[..., 0] != sh.UNEVALUATED).sum())
        done = sum(len(c[0]) for c in calls)
        assert done <= 2.2 * on_path + 8, (done, on_path)  # probes + guesses cost about as much again as the true search path


GLOO_BATCHED_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["VFSMS_ROOT"])
import torch.distributed as dist
from imagestitch_b200 import sharding as sh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
true_dir = [1] * 5 + [2] + [3] * 5 + [2] + [1] * 5 + [2] + [3] * 4          # 2 ranks: the second shard starts at a turn
n = len(true_dir)
ranges = sh.partition_pairs(n, world)
lo, hi = ranges[rank]
def evaluate(pair, i, d):
    return int(d == true_dir[pair] and i >= 1 + (pair % 5 == 2)), 50 + pair, pair - 3, 7
def batch_evaluate(pairs, i, dirs):
    assert all(lo <= p < hi for p in pairs), (rank, pairs)          # only the owner is ever asked
    return [evaluate(p, i, d) for p, d in zip(pairs, dirs)]
seq, _ = sh.replay(sh.evaluate_shard(evaluate, 0, n, 1, 1, 0.2), evaluate, 1, 1, 0.2)
out, stats = sh.align_sequence_sharded_batched(batch_evaluate, n, 1, 1, 0.2, rank, world)               # with direction probes
assert out == seq, (rank, out, seq)
out, stats = sh.align_sequence_sharded_batched(batch_evaluate, n, 1, 1, 0.2, rank, world, probe_step=0)   # blind walk from direction 1
assert out == seq, (rank, out, seq)
if rank == 0:
    assert stats["extra_rounds"] >= 1 and stats["on_demand"] >= 2          # pair 11: carried direction 3, the owner assumed 1
    print("GLOO_BATCHED_OK", len(out), stats["extra_rounds"], stats["on_demand"])
dist.destroy_process_group()
"""


def test_sharding_batched_two_ranks_gloo(tmp_path):
    """Tiles never leave their owner: misses of the replay go back to the owning rank, results equal the sequential loop."""
    script = tmp_path / "worker_batched.py"
    script.write_text(GLOO_BATCHED_WORKER)
    env = dict(os.environ, VFSMS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_BATCHED_OK 22" in r.stdout


GLOO_BATCHED_WORKER_4 = r"""
import os, sys, threading
import numpy as np
sys.path.insert(0, os.environ["VFSMS_ROOT"])
import torch.distributed as dist
from imagestitch_b200 import sharding as sh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# a 6 x 7 serpentine: runs of 6 pairs along a row (directions 2 / 4 alternating), one pair down (direction 1) at every row end
true_dir = []
for r in range(6):
    true_dir += [2 if r % 2 == 0 else 4] * 6 + [1]
true_dir = true_dir[:-1]
n = len(true_dir)
ranges = sh.partition_pairs(n, world)
lo, hi = ranges[rank]
def evaluate(pair, i, d):
    return int(d == true_dir[pair] and i >= 1 + (pair % 7 == 3)), 50 + pair, pair - 3, 7
threads_seen = set()
def batch_evaluate(pairs, i, dirs):
    assert all(lo <= p < hi for p in pairs), (rank, pairs)
    assert len({d in (2, 4) for d in dirs}) == 1                     # one strip shape per call
    threads_seen.add(threading.get_ident())
    return [evaluate(p, i, d) for p, d in zip(pairs, dirs)]
def many(calls):                                                     # the two-lane form of tiles_batch_evaluator, on host threads
    out = [None] * len(calls)
    def work(lane):
        for k in range(lane, len(calls), 2):
            out[k] = batch_evaluate(*calls[k])
    t = threading.Thread(target=work, args=(1,)); t.start(); work(0); t.join()
    return out
batch_evaluate.many = many
seq, _ = sh.replay(sh.evaluate_shard(evaluate, 0, n, 1, 1, 0.2), evaluate, 1, 1, 0.2)
out, stats = sh.align_sequence_sharded_batched(batch_evaluate, n, 1, 1, 0.2, rank, world)
assert out == seq, (rank, out, seq)
assert len(threads_seen) == 2, threads_seen                          # rounds with both shapes really ran on two threads
if rank == 0:
    print("GLOO_BATCHED4_OK", len(out), world, stats["device_calls"])
dist.destroy_process_group()
"""


def test_sharding_batched_four_ranks_two_lanes_gloo(tmp_path):
    """Four shards of a serpentine, every round's two strip shapes evaluated side by side (batch_evaluate.many): the gathered
    replay equals the sequential loop."""
    script = tmp_path / "worker_batched4.py"
    script.write_text(GLOO_BATCHED_WORKER_4)
    env = dict(os.environ, VFSMS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29545", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_BATCHED4_OK 41 4" in r.stdout


def test_synthetic_generator_is_seeded():
    from imagestitch_b200 import synth
    a1, b1, o1 = synth.pair(seed=9, size=256, overlap=40, direction=1)
    a2, b2, o2 = synth.pair(seed=9, size=256, overlap=40, direction=1)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and o1 == o2
    assert a1.dtype == np.uint8 and 90 < a1.mean() < 160 and a1.std() > 20
    # the overlap really is the same scene
    ov = 256 - o1[0]
    assert np.corrcoef(a1[256 - ov:, 8:200].ravel().astype(float), b1[:ov, 8 - o1[1]:200 - o1[1]].ravel().astype(float))[0, 1] > 0.9


def test_get_stitch_by_offset_bookkeeping_without_gpu(tmp_path, golden_dir, monkeypatch):
    """Stitcher.getStitchByOffset's host half (Stitcher.py:378-431): origins / ROI rectangles / canvas size handed to the device
    mosaic equal the oracle's closed form, and rendering those arguments with the NumPy oracle reproduces the reference's mosaic."""
    import cv2
    from imagestitch_b200 import gpu
    from imagestitch_b200.Stitcher import Stitcher
    from oracle import blend_oracle as bo
    m = np.load(os.path.join(golden_dir, "mosaic_case.npz"))
    files = []
    for k, t in enumerate(m["tiles"]):
        f = str(tmp_path / ("t%02d.png" % k)); cv2.imwrite(f, t); files.append(f)
    seen = {}

    def fake_mosaic(tiles, tile_origin, roi_rect, pair_offset, method, canvas_shape, device=0):
        seen.update(origin=np.asarray(tile_origin), roi=np.asarray(roi_rect), pair=np.asarray(pair_offset), shape=tuple(canvas_shape))
        out, _ = bo.band_renderer()(tiles, tile_origin, roi_rect, pair_offset, method, canvas_shape, False, None, None, None)
        return out
    monkeypatch.setattr(gpu, "mosaic", fake_mosaic)
    st = Stitcher()
    monkeypatch.setattr(Stitcher, "isPrintLog", False)
    monkeypatch.setattr(Stitcher, "fuseMethod", "fadeInAndFadeOut")
    monkeypatch.setattr(Stitcher, "isColorMode", False)
    monkeypatch.setattr(Stitcher, "decoder", "cv2")
    offsets = [list(map(int, o)) for o in m["offsets"]]
    out = st.getStitchByOffset(files, offsets)
    org, rois, shape = bo.rectify(m["offsets"], m["tiles"].shape[1:3])
    assert np.array_equal(seen["origin"], org) and np.array_equal(seen["roi"], rois) and seen["shape"] == shape
    assert offsets[0] == [0, 0] and len(offsets) == len(m["offsets"]) + 1      # the reference mutates its argument (Stitcher.py:386)
    assert np.array_equal(out, m["out_fadeInAndFadeOut_gray"])
