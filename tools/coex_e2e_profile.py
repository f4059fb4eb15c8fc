#!/usr/bin/env python3
"""Host-side profile of one end-to-end coex call at the headline shape (pinned host input, pinned outputs)."""
import cProfile
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
from normalisr_b200 import synth
from normalisr_b200 import normalisr as norm

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device("cuda", 0)
p = synth.device_problem(1001, genes, 100000, dev)
dt = torch.empty(p["dt"].shape, dtype=torch.float64, pin_memory=True).copy_(p["dt"])
dc = p["dc"].cpu()
del p
torch.cuda.empty_cache()
P = torch.empty((genes, genes), dtype=torch.float64, pin_memory=True)
D = torch.empty((genes, genes), dtype=torch.float64, pin_memory=True)
for _ in range(2):
    norm.coex(dt, dc, out=(P, D))
t0 = time.perf_counter()
for _ in range(3):
    norm.coex(dt, dc, out=(P, D))
print("ms per call", (time.perf_counter() - t0) / 3 * 1e3)
from normalisr_b200 import association
for tiles, chunk_bytes in ((12, 1 << 30), (4, 1 << 29), (2, 1 << 28)):
    association._STRIP_TILES, association._PIPE_CHUNK_BYTES = tiles, chunk_bytes
    norm.coex(dt, dc, out=(P, D))
    t0 = time.perf_counter()
    for _ in range(3):
        norm.coex(dt, dc, out=(P, D))
    print("strip tiles %2d, chunk bytes 2^%d: %.1f ms per call" % (tiles, chunk_bytes.bit_length() - 1, (time.perf_counter() - t0) / 3 * 1e3), flush=True)
association._STRIP_TILES, association._PIPE_CHUNK_BYTES = 2, 1 << 28
association._TIMELINE = []
t_host0 = time.perf_counter()
e_start = torch.cuda.Event(enable_timing=True)
e_start.record()
norm.coex(dt, dc, out=(P, D))
torch.cuda.synchronize()
print("timed call ms", (time.perf_counter() - t_host0) * 1e3)
rows = {}
for label, chunk, ev in association._TIMELINE:
    rows.setdefault(chunk, {})[label] = e_start.elapsed_time(ev)
print("chunk  h2d_begin  h2d_end  projected  contracted  d2h_end   (ms since the call started)")
for c in sorted(rows):
    r = rows[c]
    print("%3d  %8.1f %8.1f %8.1f %8.1f %8.1f" % (c, r.get("h2d_begin", -1), r.get("h2d_end", -1), r.get("projected", -1),
                                                   r.get("contracted", -1), r.get("d2h_end", -1)))
association._TIMELINE = None
pr = cProfile.Profile()
pr.enable()
for _ in range(2):
    norm.coex(dt, dc, out=(P, D))
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
