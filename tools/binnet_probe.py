#!/usr/bin/env python3
"""binnet on the P of a 20,000-gene co-expression run (cells as given): times the row kernels; under ncu
(-k regex:binnet_rows) it is the profiling target."""
import sys
import torch
sys.path.insert(0, ".")
from normalisr_b200 import binnet as bn, engine, synth
from normalisr_b200 import normalisr as norm

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
torch.cuda.set_device(0)
ctx = engine.context(0)
p = synth.device_problem(1001, 20000, cells, torch.device("cuda", 0))
P, _, _ = norm.coex(p["dt"], p["dc"])
del p
out = torch.empty(P.shape, dtype=torch.uint8, device=P.device)
stats = torch.zeros(2, dtype=torch.int64, device=P.device)
for keys in (1, 0):
    engine.set_option("binnet_keys", keys)
    for q in (0.05, 1e-4):
        for _ in range(2):
            bn.binnet_rows(ctx, P, q, 0, out=out, stats=stats)
        stats.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            bn.binnet_rows(ctx, P, q, 0, out=out, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        print("keys=%d qcut=%g: %.3f ms, edges %d" % (keys, q, e0.elapsed_time(e1) / 5, int(stats[0]) // 5), flush=True)
engine.set_option("binnet_keys", 1)
