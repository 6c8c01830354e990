"""Per-call cost of vfsms_tiles_align_list for small batches on one B200 (what a round of the sharded search costs):
wall time per call and the describe / Hessian / matcher stage times for 1, 4, 8, 12 and 16 ROI pairs.
  gpurun -- python scripts/diag_call_latency.py"""
import time, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from imagestitch_b200 import gpu, synth
dev = torch.device("cuda:0")
TILE, OVERLAP = 2048, 205
tiles, off = synth.sequence_torch(4242, 2, 8, TILE, OVERLAP, dev, first=0, count=16)
gpu.tiles_reserve(16, TILE, TILE)
gpu.tiles_upload(0, tiles.cpu().numpy())
params = gpu.surf_params()
L = int(0.2 * TILE)
gpu.profile_enable(True)
for n, dirs in ((1, [2]), (4, [2, 4, 2, 4]), (8, [2] * 8), (12, [2] * 12), (16, [2] * 16)):
    firsts = [f % 12 for f in range(n)]
    for _ in range(2):
        gpu.tiles_align_list(firsts, dirs, L, params=params)
    gpu.profile_read(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r = gpu.tiles_align_list(firsts, dirs, L, params=params)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5 * 1e3
    st = gpu.profile_read(True)
    print("evals %2d: %.2f ms per call, describe %.2f hessian %.2f match %.2f" % (n, dt, st["orient_describe"][0] / 5, st["hessian_nms"][0] / 5, st["match_tc"][0] / 5))
