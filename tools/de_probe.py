#!/usr/bin/env python3
"""One warm pass of the DE workloads (for `ncu --metrics gpu__time_duration.sum` launch lists):
config 3 (50k cells x 10k genes x 300 gRNAs, single = 0 / 4 / 1) and the per-GPU share of config 5
(1M cells x 2,500 genes x 1,000 gRNAs, single = 0).  usage: de_probe.py [c3|c5] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from normalisr_b200 import normalisr as norm, parallel, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
if which == "c3":
    p = synth.device_problem(1003, 10000, 50000, dev, n_group=300, group_p=0.02)
    for single in (0, 4):
        for _ in range(reps):
            torch.cuda.nvtx.range_push("de_single%d" % single)
            norm.de(p["dg"], p["dt"], p["dc"], single=single)
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
else:
    p = synth.device_problem(1005, 2500, 1000000, dev, n_group=1000, group_p=0.002, n_module=0)
    for _ in range(reps):
        parallel.de_sharded(p["dg"], p["dt"], p["dc"], 2500)
        torch.cuda.synchronize()
print("done", which)
