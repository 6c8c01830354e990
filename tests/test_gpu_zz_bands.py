"""GPU: one band of a partitioned mosaic (vfsms_mosaic_band_host) and the whole band chain (sharding.mosaic_sharded's
building blocks) against the single-call device mosaic and the NumPy oracle.  Single GPU: the bands run one after the
other on cuda:0, patches are handed over in process (the transport is covered by the gloo test on the CPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return g


def _chain(tiles, offs, method, world):
    from imagestitch_b200 import sharding as sh
    h, w = tiles.shape[1:3]
    full = [[0, 0]] + offs
    origins, rois, shape = sh.rectify_offsets(full, [(h, w)] * len(full))
    ranges, boxes = sh.plan_mosaic_bands(origins, (h, w), world)
    active = [r for r in range(world) if ranges[r][1] > ranges[r][0]]
    render = sh.gpu_band_renderer(0)
    patches, bands = [], []
    for pos, r in enumerate(active):
        s0, s1 = ranges[r]
        canvas, patches = sh.render_band(render, tiles[s0:s1], origins, rois, full, method, (s0, s1), boxes[r], patches,
                                         [boxes[q] for q in active[pos + 1:]])
        bands.append(((s0, s1), boxes[r], canvas))
    return sh.compose_bands(shape, tiles.shape[3:], bands, origins, (h, w)), (origins, rois, full, shape)


@pytest.mark.parametrize("method", ["fadeInAndFadeOut", "average", "notFuse"])
def test_band_chain_equals_single_call_and_oracle(gpu, method):
    from oracle import blend_oracle as bo
    from test_mosaic_bands_cpu import serpentine, tiles_for
    rng = np.random.default_rng(8)
    n, cols, h, w = 17, 4, 96, 128
    offs = serpentine(rng, n, cols, h, w, 14, 18, 3)
    tiles = tiles_for(rng, n, h, w, False)
    ref = bo.mosaic(tiles, offs, method)
    for world in (1, 2, 3, 8):
        out, (origins, rois, full, shape) = _chain(tiles, offs, method, world)
        assert np.array_equal(out, ref), (method, world)
    single = gpu.mosaic(tiles, origins, rois, np.asarray(full, np.int32), method, shape)
    assert np.array_equal(single, ref)


def test_band_chain_colour(gpu):
    from oracle import blend_oracle as bo
    from test_mosaic_bands_cpu import serpentine, tiles_for
    rng = np.random.default_rng(9)
    offs = serpentine(rng, 9, 3, 64, 80, 10, 12, 2)
    tiles = tiles_for(rng, 9, 64, 80, True)
    ref = bo.mosaic(tiles, offs, "fadeInAndFadeOut")
    for world in (2, 4):
        assert np.array_equal(_chain(tiles, offs, "fadeInAndFadeOut", world)[0], ref)


def test_band_halo_roundtrip_and_errors(gpu):
    """halo_in is pasted verbatim (holes stay holes), halo_out returns int16 with -1 where nothing was written."""
    from imagestitch_b200 import _lib
    tile = np.full((1, 8, 8), 50, np.uint8)
    halo = np.full((4, 6), -1, np.int16); halo[1:3, 2:5] = 7
    out, back = gpu.mosaic_band(tile, [[10, 10]], [[0, 0, 0, 0]], [[0, 0]], "notFuse", (20, 24), fuse_first=True,
                                halo_in=halo, halo_in_rect=(0, 0, 4, 6), halo_out_rect=(0, 0, 20, 24))
    exp = np.full((20, 24), -1, np.int16); exp[0:4, 0:6] = halo; exp[10:18, 10:18] = 50
    assert np.array_equal(back, exp)
    assert np.array_equal(out, np.where(exp < 0, 0, exp).astype(np.uint8))
    with pytest.raises(_lib.VfsmsError):
        gpu.mosaic_band(tile, [[10, 10]], [[0, 0, 0, 0]], [[0, 0]], "notFuse", (20, 24), halo_in=halo, halo_in_rect=(18, 0, 4, 6))
