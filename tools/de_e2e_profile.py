#!/usr/bin/env python3
"""Host-side profile of one end-to-end de(single=4) call at the config-3 shape (pinned host inputs)."""
import cProfile
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
from normalisr_b200 import synth
from normalisr_b200 import normalisr as norm

dev = torch.device("cuda", 0)
p = synth.device_problem(1002, 10000, 50000, dev, n_group=300, group_p=0.02, n_module=0)
hosts = {k: torch.empty(p[k].shape, dtype=torch.float64, pin_memory=True).copy_(p[k]) for k in ("dg", "dt", "dc")}
del p
torch.cuda.empty_cache()
for _ in range(2):
    norm.de(hosts["dg"], hosts["dt"], hosts["dc"], single=4)
t0 = time.perf_counter()
for _ in range(3):
    norm.de(hosts["dg"], hosts["dt"], hosts["dc"], single=4)
print("ms per call", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    norm.de(hosts["dg"], hosts["dt"], hosts["dc"], single=4)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
