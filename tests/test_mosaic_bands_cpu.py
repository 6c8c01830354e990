"""CPU: the band partition of the mosaic loop (sharding.mosaic_sharded) reproduces the sequential
loop byte for byte for any number of ranks -- with the NumPy oracle as the renderer, in-process for every world size
and through torch.distributed (gloo, 3 ranks) for the transport."""
import os
import subprocess
import sys

import numpy as np
import pytest

from imagestitch_b200 import sharding as sh
from oracle import blend_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def serpentine(rng, n, cols, h, w, ov_r, ov_c, jitter):
    offs = []
    for k in range(1, n):
        j = [int(v) for v in rng.integers(-jitter, jitter + 1, 2)]
        if k % cols == 0:
            offs.append([h - ov_r + j[0], j[1]])
        else:
            s = 1 if (k // cols) % 2 == 0 else -1
            offs.append([j[0], s * (w - ov_c + j[1])])
    return offs


def tiles_for(rng, n, h, w, color):
    t = rng.integers(1, 256, (n, h, w) + ((3,) if color else ())).astype(np.uint8)
    t[:, ::7, ::5] = 0                         # real black pixels: treated as holes by the non-fade modes (Appendix C)
    return t


def chain_in_process(tiles, offs, method, world):
    """mosaic_sharded without a process group: the same plan / render_band / compose calls, patches handed over in order."""
    h, w = tiles.shape[1:3]
    full = [[0, 0]] + offs
    origins, rois, shape = sh.rectify_offsets(full, [(h, w)] * len(full))
    ranges, boxes = sh.plan_mosaic_bands(origins, (h, w), world)
    active = [r for r in range(world) if ranges[r][1] > ranges[r][0]]
    patches, bands = [], []
    for pos, r in enumerate(active):
        s0, s1 = ranges[r]
        canvas, patches = sh.render_band(bo.band_renderer(), tiles[s0:s1], origins, rois, full, method, (s0, s1), boxes[r],
                                         patches, [boxes[q] for q in active[pos + 1:]])
        bands.append(((s0, s1), boxes[r], canvas))
    return sh.compose_bands(shape, tiles.shape[3:], bands, origins, (h, w)), patches


@pytest.mark.parametrize("method", ["fadeInAndFadeOut", "average", "trigonometric", "notFuse"])
def test_bands_equal_sequential_serpentine(method):
    rng = np.random.default_rng(3)
    n, cols, h, w = 23, 5, 40, 52
    offs = serpentine(rng, n, cols, h, w, 9, 11, 3)
    tiles = tiles_for(rng, n, h, w, False)
    ref = bo.mosaic(tiles, offs, method)
    for world in (1, 2, 3, 4, 5, 8, 23, 30):       # band cuts in the middle of grid rows, one tile per rank, idle ranks
        out, left = chain_in_process(tiles, offs, method, world)
        assert np.array_equal(out, ref), (method, world)
        assert left == []                          # nothing is forwarded past the last band


def test_bands_equal_sequential_colour_and_random_walk():
    rng = np.random.default_rng(11)
    # colour, fade
    offs = serpentine(rng, 12, 4, 30, 34, 8, 9, 2)
    tiles = tiles_for(rng, 12, 30, 34, True)
    ref = bo.mosaic(tiles, offs, "fadeInAndFadeOut")
    for world in (2, 3, 7):
        assert np.array_equal(chain_in_process(tiles, offs, "fadeInAndFadeOut", world)[0], ref)
    # arbitrary walks: tiles of non-adjacent bands overlap, origins shift (negative running sums), patches are forwarded
    # through bands that do not touch them
    for trial in range(12):
        n = int(rng.integers(2, 14))
        h, w = 24, 28
        offs = [[int(v) for v in rng.integers(-20, 21, 2)] for _ in range(n - 1)]
        tiles = tiles_for(rng, n, h, w, False)
        for method in ("fadeInAndFadeOut", "maximum"):
            try:
                ref = bo.mosaic(tiles, offs, method)
            except (IndexError, ZeroDivisionError):
                continue                            # the reference itself raises on this geometry (getWeightsMatrix quirks)
            for world in (2, 3, n):
                assert np.array_equal(chain_in_process(tiles, offs, method, world)[0], ref), (trial, method, world)


def test_band_plan_properties():
    origins, rois, shape = sh.rectify_offsets([[0, 0], [0, 90], [0, 90], [80, 0], [0, -90], [0, -90]], [(100, 100)] * 6)
    ranges, boxes = sh.plan_mosaic_bands(origins, (100, 100), 4)
    assert ranges == [(0, 2), (2, 4), (4, 5), (5, 6)]
    assert boxes[0] == (0, 0, 100, 190) and boxes[1] == (0, 180, 180, 280)
    ranges, boxes = sh.plan_mosaic_bands(origins[:2], (100, 100), 4)
    assert [b is None for b in boxes] == [False, False, True, True]


GLOO_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["VFSMS_ROOT"])
import torch.distributed as dist
from imagestitch_b200 import sharding as sh
from oracle import blend_oracle as bo
sys.path.insert(0, os.path.join(os.environ["VFSMS_ROOT"], "tests"))
from test_mosaic_bands_cpu import serpentine, tiles_for
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(21)
n, cols, h, w = 14, 4, 36, 44
offs = serpentine(rng, n, cols, h, w, 8, 10, 2)
tiles = tiles_for(rng, n, h, w, False)
loaded = []
def load(s, e):
    loaded.append((s, e))
    return tiles[s:e]
out = sh.mosaic_sharded(bo.band_renderer(), load, offs, (h, w), "fadeInAndFadeOut", rank, world)
assert loaded == [sh.partition_pairs(n, world)[rank]], loaded        # every rank touches only its own tiles
if rank == 0:
    ref = bo.mosaic(tiles, offs, "fadeInAndFadeOut")
    assert out.shape == ref.shape and np.array_equal(out, ref)
    print("BANDS_OK", out.shape)
dist.barrier()
dist.destroy_process_group()
'''


def test_mosaic_bands_three_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, VFSMS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
                        "--master-port", "29547", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BANDS_OK" in r.stdout
