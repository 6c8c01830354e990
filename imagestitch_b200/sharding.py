"""Multi-GPU sharding of a tile sequence's pairwise alignments.

The unit of work is the consecutive pair (k, k+1); pairs are independent except for one carried integer, the search
direction (Stitcher.py:317,361).  One process per GPU:
  1. pairs are partitioned contiguously over ranks (each tile is touched by at most two ranks);
  2. every rank walks ITS pairs in order with the reference's search loop, assuming the global initial direction for its
     first pair, and records every candidate it evaluates in a table  T[pair][i][direction] = (status, dRow, dCol, votes);
  3. ONE collective: all_gather of the int32 tables (NCCL over NVLink on GPUs, gloo in the CPU tests) -- KBs;
  4. rank 0 replays the reference's sequential loop over the gathered table; a candidate the owning rank did not
     evaluate (its direction guess at the shard boundary was wrong AND the true order visits something else first) is
     evaluated on demand.  Candidate results are pure functions of (pair, i, direction), so the replayed offsets are
     exactly those of the sequential loop.
No other data crosses GPUs.  `evaluate(pair, i, direction)` is injected: on GPUs it is the fused align call, in the CPU
tests a deterministic fake.
"""
import numpy as np

UNEVALUATED = -(2 ** 31)


def partition_pairs(n_pairs, world_size):
    """Contiguous [start, stop) ranges, sizes differing by at most one."""
    base, extra = divmod(n_pairs, world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def max_i(roi_ratio):
    """Number of ROI sizes tried: i in 1 .. maxI-1 (Stitcher.py:316)."""
    return int(np.floor(0.5 / roi_ratio) + 1) + 1


def direction_increase(direction, direct_incre):
    """Stitcher.directionIncrease (Stitcher.py:36-47)."""
    direction += direct_incre
    if direction == 5:
        direction = 1
    if direction == 0:
        direction = 4
    return direction


def search_pair(evaluate, pair, direction, direct_incre, roi_ratio, table=None):
    """The reference's candidate order for one pair (Stitcher.py:319-351).  Returns (status, i, direction, (dRow, dCol)).
    `table` (int32 [n_i, 4, 4] for this pair) caches / records evaluations."""
    ini = direction
    local = ini
    for i in range(1, max_i(roi_ratio)):
        while True:
            if table is not None and table[i - 1, local - 1, 0] != UNEVALUATED:
                st, dr, dc, votes = (int(v) for v in table[i - 1, local - 1])
            else:
                st, dr, dc, votes = evaluate(pair, i, local)
                if table is not None:
                    table[i - 1, local - 1] = (int(st), dr, dc, votes)
            if st:
                return True, i, local, (dr, dc)
            local = direction_increase(local, direct_incre)
            if local == ini:
                break
    return False, 0, ini, (0, 0)


def evaluate_shard(evaluate, start, stop, direction, direct_incre, roi_ratio):
    """Step 2: walk pairs [start, stop) sequentially.  -> int32 table [stop-start, n_i, 4, 4]."""
    n_i = max_i(roi_ratio) - 1
    table = np.full((stop - start, n_i, 4, 4), UNEVALUATED, np.int32)
    d = direction
    for k in range(start, stop):
        st, _, d_new, _ = search_pair(evaluate, k, d, direct_incre, roi_ratio, table[k - start])
        if st:
            d = d_new          # a failed pair ends the segment; the next segment starts with the same carried direction
    return table


def replay(table, evaluate, direction, direct_incre, roi_ratio):
    """Step 4: the sequential loop over the gathered table.  -> list of (status, i, direction, (dRow, dCol)) per pair
    and the number of candidates that had to be evaluated on demand."""
    out, on_demand = [], 0
    d = direction
    for k in range(table.shape[0]):
        before = int((table[k, :, :, 0] != UNEVALUATED).sum())
        st, i, d_new, off = search_pair(evaluate, k, d, direct_incre, roi_ratio, table[k])
        on_demand += int((table[k, :, :, 0] != UNEVALUATED).sum()) - before
        out.append((st, i, d_new, off))
        if st:
            d = d_new
    return out, on_demand


def gather_tables(local_table, ranges, rank, world_size, device=None):
    """Step 3: one all_gather of equally padded int32 tables.  Works on any initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist
    n_max = max(e - s for s, e in ranges)
    shape = (n_max,) + tuple(local_table.shape[1:])
    buf = torch.full(shape, UNEVALUATED, dtype=torch.int32)
    buf[: local_table.shape[0]] = torch.from_numpy(local_table)
    if device is not None:
        buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(parts, buf)
    full = [p.cpu().numpy()[: e - s] for p, (s, e) in zip(parts, ranges)]
    return np.concatenate(full, axis=0)


def align_sequence_sharded(evaluate, n_pairs, direction, direct_incre, roi_ratio, rank, world_size, device=None):
    """Steps 1-4.  Every rank returns the same replayed result list (rank 0's replay is what a caller should use)."""
    ranges = partition_pairs(n_pairs, world_size)
    s, e = ranges[rank]
    local = evaluate_shard(evaluate, s, e, direction, direct_incre, roi_ratio)
    full = gather_tables(local, ranges, rank, world_size, device) if world_size > 1 else local
    return replay(full, evaluate, direction, direct_incre, roi_ratio)


def roi_origin_back(offset, shape_a, shape_b, i, direction, roi_ratio):
    """Add the ROI origin back (Stitcher.py:353-360)."""
    off = [int(offset[0]), int(offset[1])]
    if direction == 1:
        off[0] += shape_a[0] - int(i * roi_ratio * shape_a[0])
    elif direction == 2:
        off[1] += shape_a[1] - int(i * roi_ratio * shape_a[1])
    elif direction == 3:
        off[0] -= shape_b[0] - int(i * roi_ratio * shape_b[0])
    elif direction == 4:
        off[1] -= shape_b[1] - int(i * roi_ratio * shape_b[1])
    return off


def gpu_evaluator(tiles, params=None, ratio=0.75, offset_evaluate=3, roi_ratio=0.2, device=0):
    """evaluate(pair, i, direction) backed by the fused device call on host tiles [n, H, W]."""
    from . import gpu
    from .ImageUtility import Method
    m = Method()

    def evaluate(pair, i, direction):
        a = m.getROIRegionForIncreMethod(tiles[pair], direction, "first", i * roi_ratio)
        b = m.getROIRegionForIncreMethod(tiles[pair + 1], direction, "second", i * roi_ratio)
        r = gpu.align_batch(a[None], b[None], params=params, ratio=ratio, offset_evaluate=offset_evaluate, device=device)[0]
        return int(r["status"]), int(r["d_row"]), int(r["d_col"]), int(r["votes"])
    return evaluate


# ---------------------------------------------------------------- mosaic: tile sequence partitioned into bands
# The reference's paste/blend loop (Stitcher.py:440-483) is sequential: tile i blends against everything pasted before it, and
# the corner weights (ImageFusion.py:43-190) depend on WHICH pixels of its ROI are already filled.  Partition:
#   1. every rank runs the integer bookkeeping for the whole sequence (rectify_offsets: tile origins, ROI rectangles, canvas);
#   2. tiles are split into contiguous runs ("bands", one per rank); a band's canvas is the bounding box of its tile rectangles;
#   3. ranks render in chain order: rank s receives the frontier -- int16 patches (holes = -1) of what earlier bands left
#      inside later bands' boxes --, pastes the part inside its own box, renders its tiles in reference order, and forwards
#      the surviving old patches plus one new patch (its canvas clipped to the later boxes);
#      loading / decoding / uploading of the tiles -- the bulk of the time -- runs concurrently on all ranks, only the
#      render itself waits for the patch (KBs to MBs) from the previous rank;
#   4. optional: rank 0 gathers the band canvases and paints each band's OWN tile rectangles in rank order (a later band's
#      pixels supersede an earlier band's, exactly like the sequential loop's overwrites).
# Every pixel a tile's ROI reads is either written by a tile of the same band or arrives in a patch, so the output is
# byte-identical to the sequential loop for any partition.  `render` and the point-to-point transport are injected:
# gpu.mosaic_band + torch.distributed on GPUs, a NumPy renderer + gloo in the CPU tests.

def rectify_offsets(origin_offsets, tile_shapes):
    """Global offset rectification, Stitcher.getStitchByOffset's integer bookkeeping (Stitcher.py:386-431).
    origin_offsets: [[dRow, dCol], ...] per tile, entry 0 = [0, 0]; tile_shapes: [(rows, cols), ...].
    -> (origins int32 [n, 2], rois int32 [n, 4] (r0, c0, r1, c1; row 0 unused), (canvas_rows, canvas_cols))."""
    n = len(origin_offsets)
    off = [[int(o[0]), int(o[1])] for o in origin_offsets]
    range_x = [[0, 0] for _ in range(n)]
    range_y = [[0, 0] for _ in range(n)]
    result_row, result_col = int(tile_shapes[0][0]), int(tile_shapes[0][1])
    range_x[0][1], range_y[0][1] = result_row, result_col
    dx_sum = dy_sum = 0
    for i in range(1, n):
        h, w = int(tile_shapes[i][0]), int(tile_shapes[i][1])
        dx_sum += off[i][0]
        dy_sum += off[i][1]
        if dx_sum <= 0:
            for j in range(i):
                off[j][0] += abs(dx_sum)
                range_x[j][0] += abs(dx_sum)
                range_x[j][1] += abs(dx_sum)
            result_row += abs(dx_sum)
            range_x[i][1] = result_row
            dx_sum = range_x[i][0] = off[i][0] = 0
        else:
            off[i][0] = dx_sum
            result_row = max(result_row, dx_sum + h)
            range_x[i][1] = result_row
        if dy_sum <= 0:
            for j in range(i):
                off[j][1] += abs(dy_sum)
                range_y[j][0] += abs(dy_sum)
                range_y[j][1] += abs(dy_sum)
            result_col += abs(dy_sum)
            range_y[i][1] = result_col
            dy_sum = range_y[i][0] = off[i][1] = 0
        else:
            off[i][1] = dy_sum
            result_col = max(result_col, dy_sum + w)
            range_y[i][1] = result_col
    rois = np.zeros((n, 4), np.int32)
    for i in range(1, n):
        h, w = int(tile_shapes[i][0]), int(tile_shapes[i][1])
        rois[i] = (max(off[i][0], range_x[i - 1][0]), max(off[i][1], range_y[i - 1][0]),
                   min(off[i][0] + h, range_x[i - 1][1]), min(off[i][1] + w, range_y[i - 1][1]))
    return np.asarray(off, np.int32).reshape(n, 2), rois, (result_row, result_col)


def _intersect(a, b):
    r = (max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3]))
    return r if r[0] < r[2] and r[1] < r[3] else None


def _bounding(rects):
    rects = [r for r in rects if r is not None]
    if not rects:
        return None
    return (min(r[0] for r in rects), min(r[1] for r in rects), max(r[2] for r in rects), max(r[3] for r in rects))


def plan_mosaic_bands(origins, tile_shape, world_size):
    """-> (ranges [(start, stop)], boxes [(r0, c0, r1, c1) or None]) : contiguous tile runs and their bounding boxes."""
    n = len(origins)
    h, w = int(tile_shape[0]), int(tile_shape[1])
    ranges = partition_pairs(n, world_size)
    boxes = []
    for s, e in ranges:
        boxes.append(_bounding([(int(origins[i][0]), int(origins[i][1]), int(origins[i][0]) + h, int(origins[i][1]) + w)
                                for i in range(s, e)]))
    return ranges, boxes


def render_band(render, tiles, origins, rois, pair_offsets, method, rng, box, patches_in, later_boxes):
    """Render the tiles [rng) on the canvas `box` given the frontier `patches_in` = [((r0, c0, r1, c1), int16 array), ...]
    (global coordinates, oldest first).  -> (canvas uint8, patches_out)."""
    s0, s1 = rng
    r0, c0 = box[0], box[1]
    shape = (box[2] - box[0], box[3] - box[1])
    chan = tuple(np.asarray(tiles).shape[3:])
    local_org = np.asarray(origins[s0:s1], np.int32) - np.asarray([r0, c0], np.int32)
    local_roi = np.asarray(rois[s0:s1], np.int32) - np.asarray([r0, c0, r0, c0], np.int32)
    if s0 == 0:
        local_roi[0] = 0                      # tile 0 of the sequence has no ROI
    halo_in = halo_in_rect = None
    inside = [_intersect(rect, box) for rect, _ in patches_in]
    hb = _bounding(inside)
    if hb is not None:
        halo_in = np.full((hb[2] - hb[0], hb[3] - hb[1]) + chan, -1, np.int16)
        for (rect, data), cut in zip(patches_in, inside):
            if cut is None:
                continue
            src = data[cut[0] - rect[0]:cut[2] - rect[0], cut[1] - rect[1]:cut[3] - rect[1]]
            dst = halo_in[cut[0] - hb[0]:cut[2] - hb[0], cut[1] - hb[1]:cut[3] - hb[1]]
            filled = src != -1                 # a newer patch never un-fills what an older one carried
            dst[filled] = src[filled]
        halo_in_rect = (hb[0] - r0, hb[1] - c0, hb[2] - hb[0], hb[3] - hb[1])
    ob = _bounding([_intersect(box, lb) for lb in later_boxes if lb is not None])
    halo_out_rect = None if ob is None else (ob[0] - r0, ob[1] - c0, ob[2] - ob[0], ob[3] - ob[1])
    canvas, halo_out = render(tiles, local_org, local_roi, np.asarray(pair_offsets[s0:s1], np.int32), method, shape,
                              fuse_first=s0 > 0, halo_in=halo_in, halo_in_rect=halo_in_rect, halo_out_rect=halo_out_rect)
    patches_out = [(rect, data) for rect, data in patches_in
                   if any(lb is not None and _intersect(rect, lb) is not None for lb in later_boxes)]
    if ob is not None:
        patches_out.append((ob, np.ascontiguousarray(halo_out, np.int16)))
    return canvas, patches_out


def compose_bands(canvas_shape, channels, bands, origins, tile_shape):
    """Step 4: bands = [((start, stop), box, canvas uint8)] in rank order -> the full mosaic."""
    h, w = int(tile_shape[0]), int(tile_shape[1])
    out = np.zeros(tuple(canvas_shape) + tuple(channels), np.uint8)
    for (s0, s1), box, canvas in bands:
        for i in range(s0, s1):
            a, b = int(origins[i][0]), int(origins[i][1])
            out[a:a + h, b:b + w] = canvas[a - box[0]:a - box[0] + h, b - box[1]:b - box[1] + w]
    return out


def _send_patches(patches, dst, device):
    import torch
    import torch.distributed as dist
    head = torch.zeros(1 + 6 * len(patches), dtype=torch.int64)
    head[0] = len(patches)
    for k, (rect, data) in enumerate(patches):
        head[1 + 6 * k:7 + 6 * k] = torch.tensor(list(rect) + [data.ndim, data.shape[2] if data.ndim == 3 else 1])
    n_head = torch.tensor([head.numel()], dtype=torch.int64)
    for t in [n_head, head] + [torch.from_numpy(np.ascontiguousarray(d, np.int16)).reshape(-1) for _, d in patches]:
        if t.numel():
            dist.send(t.to(device) if device is not None else t, dst)


def _recv_patches(src, device):
    import torch
    import torch.distributed as dist

    def recv(n, dtype):
        t = torch.empty(n, dtype=dtype, device=device) if device is not None else torch.empty(n, dtype=dtype)
        if n:
            dist.recv(t, src)
        return t.cpu()
    n_head = int(recv(1, torch.int64)[0])
    head = recv(n_head, torch.int64).tolist()
    patches = []
    for k in range(int(head[0])):
        r0, c0, r1, c1, ndim, ch = (int(v) for v in head[1 + 6 * k:7 + 6 * k])
        shape = (r1 - r0, c1 - c0) if ndim == 2 else (r1 - r0, c1 - c0, ch)
        patches.append(((r0, c0, r1, c1), recv(int(np.prod(shape)), torch.int16).numpy().reshape(shape).copy()))
    return patches


def mosaic_sharded(render, load_tiles, origin_offsets, tile_shape, method, rank, world_size, device=None, gather=True,
                   channels=()):
    """Steps 1-4.  load_tiles(start, stop) -> uint8 [stop-start, h, w(, 3)] is called once, for this rank's run only.
    origin_offsets: the pair offsets as getStitchByOffset receives them, WITHOUT the leading [0, 0].
    -> rank 0 (gather=True): the mosaic; other ranks / gather=False: ((start, stop), box, canvas uint8 of the band)."""
    import torch
    import torch.distributed as dist
    offs = [[0, 0]] + [[int(o[0]), int(o[1])] for o in origin_offsets]
    n = len(offs)
    origins, rois, canvas_shape = rectify_offsets(offs, [tile_shape] * n)
    ranges, boxes = plan_mosaic_bands(origins, tile_shape, world_size)
    active = [r for r in range(world_size) if ranges[r][1] > ranges[r][0]]     # n < world_size leaves trailing ranks idle
    band = None
    if rank in active:
        s0, s1 = ranges[rank]
        tiles = load_tiles(s0, s1)                     # concurrent on all ranks; only the render below is chained
        pos = active.index(rank)
        patches = _recv_patches(active[pos - 1], device) if pos > 0 else []
        canvas, patches = render_band(render, tiles, origins, rois, offs, method, (s0, s1), boxes[rank], patches,
                                      [boxes[r] for r in active[pos + 1:]])
        if pos + 1 < len(active):
            _send_patches(patches, active[pos + 1], device)
        band = ((s0, s1), boxes[rank], canvas)
    if not gather or world_size == 1:
        if gather:
            return compose_bands(canvas_shape, channels, [band], origins, tile_shape)
        return band
    if rank != 0:
        if band is not None:
            t = torch.from_numpy(np.ascontiguousarray(band[2])).reshape(-1)
            dist.send(t.to(device) if device is not None else t, 0)
        return band
    bands = [band]
    for r in active[1:]:
        box = boxes[r]
        shape = (box[2] - box[0], box[3] - box[1]) + tuple(channels)
        t = torch.empty(int(np.prod(shape)), dtype=torch.uint8, device=device) if device is not None \
            else torch.empty(int(np.prod(shape)), dtype=torch.uint8)
        dist.recv(t, r)
        bands.append((ranges[r], box, t.cpu().numpy().reshape(shape)))
    return compose_bands(canvas_shape, channels, bands, origins, tile_shape)


def gpu_band_renderer(device=0):
    """render(...) backed by vfsms_mosaic_band_host on this rank's GPU."""
    from . import gpu

    def render(tiles, origins, rois, pair_offsets, method, shape, fuse_first, halo_in, halo_in_rect, halo_out_rect):
        return gpu.mosaic_band(tiles, origins, rois, pair_offsets, method, shape, fuse_first=fuse_first, halo_in=halo_in,
                               halo_in_rect=halo_in_rect, halo_out_rect=halo_out_rect, device=device)
    return render
