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


# ---------------------------------------------------------------- batched variant: tiles stay on their owner's GPU
# evaluate_shard() above asks for one candidate at a time -- one device call per (pair, i, direction), the latency-bound shape.
# Here the unit of device work is a BATCH: batch_evaluate(pairs, i, directions) -> int32 [len(pairs), 4] evaluates candidate
# (i, directions[k]) of pairs[k] for a whole list in one fused call (vfsms_tiles_align_list on the rank's device-resident tiles).  A shard is walked in
# ROUNDS: the sequential search is simulated over the table filled so far, every pair stops at its first unevaluated candidate,
# and all requests of a round are grouped by (i, strip shape).  Pairs behind an unresolved pair continue with the last known
# direction -- a guess: a wrong guess only costs evaluations, the table holds pure functions of (pair, i, direction).
# Nobody but the owner can evaluate a pair (the tiles never leave its GPU), so the replay is a loop as well: rank 0 replays, the
# candidates it misses (a shard's first pairs, when the true carried direction differs from the shard's assumption) go back to
# their owners as one request list, the answers are gathered again.  Collectives: all_gather of the tables (KBs) + per extra
# round one broadcast of the request list and one all_gather of the answers.


def _missing_candidates(table, start_pair, direction, direct_incre, roi_ratio, eager=False):
    """Sequential walk over `table` (pairs start_pair ...): -> (results, requests, end_direction).  results[k] is None for a pair
    whose search stops at an unevaluated candidate; requests = [(pair, i, direction)] -- that candidate, one per such pair
    (eager: every unevaluated direction of that ROI size, which saves the rounds a turn of the path would otherwise take)."""
    results, requests = [], []
    d = direction
    for k in range(table.shape[0]):
        ini = d
        local = ini
        found = None
        pending = False
        for i in range(1, max_i(roi_ratio)):
            while True:
                st = table[k, i - 1, local - 1, 0]
                if st == UNEVALUATED:
                    requests.append((start_pair + k, i, local)); pending = True
                    if eager and direct_incre != 0:
                        requests += [(start_pair + k, i, dd) for dd in (1, 2, 3, 4) if dd != local and table[k, i - 1, dd - 1, 0] == UNEVALUATED]
                    break
                if st:
                    found = (True, i, local, (int(table[k, i - 1, local - 1, 1]), int(table[k, i - 1, local - 1, 2])))
                    break
                local = direction_increase(local, direct_incre)
                if local == ini:
                    break
            if found or pending:
                break
        if pending:
            results.append(None)
            known = np.argwhere(table[k, :, :, 0] == 1)          # the request of the pairs behind it: assume this pair ends where the
            if len(known):                                        # table already holds a success (smallest i, then direction)
                d = int(known[0][1]) + 1
        elif found:
            results.append(found); d = found[2]
        else:
            results.append((False, 0, ini, (0, 0)))
    return results, requests, d


def _run_requests(batch_evaluate, table, start_pair, requests):
    """One batch_evaluate call per (i, strip shape): directions 1 / 3 cut row strips, 2 / 4 column strips."""
    groups = {}
    for pair, i, d in requests:
        groups.setdefault((i, d in (2, 4)), []).append((pair, d))
    order = sorted(groups.items())
    calls = [([p for p, _ in items], i, [d for _, d in items]) for (i, _), items in order]
    # an evaluator with .many runs the groups of a round side by side (two strip shapes = two contexts on the GPU)
    many = getattr(batch_evaluate, "many", None)
    results = many(calls) if many is not None and len(calls) > 1 else [batch_evaluate(*c) for c in calls]
    for ((i, _), items), res in zip(order, results):
        res = np.asarray(res, np.int32).reshape(len(items), 4)
        for (pair, d), r in zip(items, res):
            table[pair - start_pair, i - 1, d - 1] = r
    return len(groups)


def evaluate_shard_batched(batch_evaluate, start, stop, direction, direct_incre, roi_ratio, probe_step=6):
    """Rounds of batched candidate evaluations over pairs [start, stop).  -> (table int32 [stop-start, n_i, 4, 4], device calls).

    A shooting path keeps its direction for long runs, but the shard does not know it (walking blindly from `direction` tests
    every direction on every pair: 4 x the work on a serpentine).  So, before the rounds: every probe_step-th pair is evaluated in
    all four directions (i = 1), every other pair once, in the direction in which its nearest probe succeeded.  The rounds then
    only fill in what the true search order also visits (the turns); speculative entries never change a result."""
    n_i = max_i(roi_ratio) - 1
    n = stop - start
    table = np.full((n, n_i, 4, 4), UNEVALUATED, np.int32)
    calls = 0
    if n > 2 and probe_step > 0 and direct_incre != 0:
        probes = list(range(start, stop, probe_step))
        calls += _run_requests(batch_evaluate, table, start, [(p, 1, d) for p in probes for d in (1, 2, 3, 4)])
        hit = {p: [d for d in (1, 2, 3, 4) if table[p - start, 0, d - 1, 0] == 1] for p in probes}
        guesses = []
        for k in range(start, stop):
            if k in hit:
                continue
            near = sorted((p for p in probes if hit[p]), key=lambda p: (abs(p - k), p > k))
            if near:
                guesses.append((k, 1, hit[near[0]][0]))
        if guesses:
            calls += _run_requests(batch_evaluate, table, start, guesses)
    while n > 0:
        _, requests, _ = _missing_candidates(table, start, direction, direct_incre, roi_ratio, eager=probe_step > 0)
        if not requests:
            break
        calls += _run_requests(batch_evaluate, table, start, requests)
    return table, calls


def align_sequence_sharded_batched(batch_evaluate, n_pairs, direction, direct_incre, roi_ratio, rank, world_size, device=None,
                                   probe_step=6):
    """Contiguous pair shards, batched rounds per shard, gather, replay on the gathered table with the misses sent back to their
    owners.  Every rank returns (results, stats): results as replay(); stats = {"device_calls", "extra_rounds", "on_demand"}."""
    import torch
    import torch.distributed as dist
    ranges = partition_pairs(n_pairs, world_size)
    s, e = ranges[rank]
    local, calls = evaluate_shard_batched(batch_evaluate, s, e, direction, direct_incre, roi_ratio, probe_step)
    extra_rounds = on_demand = 0
    while True:
        full = gather_tables(local, ranges, rank, world_size, device) if world_size > 1 else local
        # every rank replays the same gathered table, so the request list needs no broadcast
        results, requests, _ = _missing_candidates(full, 0, direction, direct_incre, roi_ratio)
        if not requests:
            break
        extra_rounds += 1
        on_demand += len(requests)
        mine = [(p, i, d) for p, i, d in requests if s <= p < e]
        if mine:
            calls += _run_requests(batch_evaluate, local, s, mine)
    return results, {"device_calls": calls, "extra_rounds": extra_rounds, "on_demand": on_demand}


def tiles_batch_evaluator(first_tile, roi_lens, params=None, ratio=0.75, offset_evaluate=3, device=0, lanes=2):
    """batch_evaluate(pairs, i, directions) on the context's device-resident tile stack: stack slot `pair - first_tile` holds tile
    `pair`.  roi_lens(i, direction) -> ROI length in pixels (int(i * roiRatio * extent), ImageUtility.py:66-101).  The directions
    of one call cut strips of one shape (sharding._run_requests groups them so).
    lanes = 2: batch_evaluate.many(calls) runs the calls of one round on two contexts of the device from two host threads (the
    second context borrows the stack, gpu.tiles_attach), so the row-strip and the column-strip candidates overlap on the GPU;
    the results are those of running the calls one after the other."""
    from . import gpu

    def batch_evaluate(pairs, i, directions, lane=0):
        r = gpu.tiles_align_list([p - first_tile for p in pairs], directions, roi_lens(i, directions[0]), params=params, ratio=ratio,
                                 offset_evaluate=offset_evaluate, device=device, lane=lane)
        return np.stack([r["status"], r["d_row"], r["d_col"], r["votes"]], axis=1).astype(np.int32)

    if lanes > 1:
        import threading

        def many(calls):
            gpu.tiles_attach(1, device=device)      # every round: the stack may have been re-reserved since (a few microseconds)
            out = [None] * len(calls)
            errors = []

            def work(lane):
                try:
                    for k in range(lane, len(calls), 2):
                        out[k] = batch_evaluate(*calls[k], lane=lane)
                except BaseException as e:          # re-raised on the calling thread
                    errors.append(e)
            t = threading.Thread(target=work, args=(1,))
            t.start()
            work(0)
            t.join()
            if errors:
                raise errors[0]
            return out
        batch_evaluate.many = many
    return batch_evaluate


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
    """Global offset rectification: the integer bookkeeping at the head of Stitcher.getStitchByOffset (Stitcher.py:386-431) in
    closed form.  origin_offsets: [[dRow, dCol], ...] per tile, entry 0 = [0, 0]; tile_shapes: [(rows, cols), ...].
    -> (origins int32 [n, 2], rois int32 [n, 4] (r0, c0, r1, c1; row 0 unused), (canvas_rows, canvas_cols)).

    Per axis, with P_i the running sum of the pair offsets: the reference shifts everything placed so far whenever the running
    position would become <= 0, which amounts to  origin_i = P_i - min(0, min_j P_j).  Its `range` bookkeeping -- the extent
    [lo_i, hi_i] of the canvas "as of tile i", later shifted like the origins -- is, in final coordinates,
        lo_i = min(0, min_{j<=i} P_j) - min(0, min_j P_j),      hi_i = max_{j<=i}(P_j + extent_j) - min(0, min_j P_j),
    except that a tile whose running position is <= 0 re-bases the axis: hi_i then only adds |P_i| to the previous extent
    (the tile's own size is NOT taken into account there -- a quirk of the reference kept as is).  The ROI of tile i is its
    rectangle clipped to the extent as of tile i - 1 (Stitcher.py:452-461).  The literal loop is restated in
    oracle/blend_oracle.py and compared with this on random and serpentine sequences (tests/test_host_cpu.py)."""
    n = len(origin_offsets)
    off = np.asarray([[int(o[0]), int(o[1])] for o in origin_offsets], np.int64).reshape(n, 2)
    ext = np.asarray([[int(t[0]), int(t[1])] for t in tile_shapes], np.int64).reshape(n, 2)
    origins = np.zeros((n, 2), np.int64)
    lo = np.zeros((n, 2), np.int64)
    hi = np.zeros((n, 2), np.int64)
    size = [0, 0]
    for ax in range(2):
        # walk once in the reference's own frame (origin re-based at every non-positive running position) ...
        pos = 0                      # running position of tile i in the current frame
        shift = 0                    # total shift applied to the frame so far
        extent = int(ext[0, ax])     # canvas extent along this axis
        frame_org = np.zeros(n, np.int64); frame_lo = np.zeros(n, np.int64); frame_hi = np.zeros(n, np.int64); at = np.zeros(n, np.int64)
        frame_hi[0] = extent
        for i in range(1, n):
            pos += int(off[i, ax])
            if pos <= 0:
                shift += -pos; extent += -pos; pos = 0
            else:
                extent = max(extent, pos + int(ext[i, ax]))
            frame_org[i] = pos; frame_lo[i] = 0; frame_hi[i] = extent; at[i] = shift
        # ... then express every record in the final frame: whatever was recorded at total shift `at[i]` moves by shift - at[i]
        origins[:, ax] = frame_org + (shift - at)
        lo[:, ax] = frame_lo + (shift - at)
        hi[:, ax] = frame_hi + (shift - at)
        size[ax] = extent
    rois = np.zeros((n, 4), np.int32)
    for i in range(1, n):
        rois[i] = (max(origins[i, 0], lo[i - 1, 0]), max(origins[i, 1], lo[i - 1, 1]),
                   min(origins[i, 0] + ext[i, 0], hi[i - 1, 0]), min(origins[i, 1] + ext[i, 1], hi[i - 1, 1]))
    return origins.astype(np.int32), rois, (int(size[0]), int(size[1]))


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
    # int16 patches travel as bytes: NCCL has no 16-bit integer type
    for t in [n_head, head] + [torch.from_numpy(np.ascontiguousarray(d, np.int16).reshape(-1).view(np.uint8)) for _, d in patches]:
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
        patches.append(((r0, c0, r1, c1), recv(2 * int(np.prod(shape)), torch.uint8).numpy().view(np.int16).reshape(shape).copy()))
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
