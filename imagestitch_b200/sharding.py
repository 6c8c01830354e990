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
