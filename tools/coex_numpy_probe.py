#!/usr/bin/env python3
"""What a caller with plain numpy arrays sees: norm.coex(dt, dc) with pageable input and freshly allocated
pageable output, against the page-locked path bench.py times."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from normalisr_b200 import synth
from normalisr_b200 import normalisr as norm

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device("cuda", 0)
p = synth.device_problem(1001, genes, 100000, dev)
dt = p["dt"].cpu().numpy()              # pageable
dc = p["dc"].cpu().numpy()
del p
torch.cuda.empty_cache()
for rep in range(3):
    t0 = time.perf_counter()
    P, D, var = norm.coex(dt, dc)
    print("numpy in / numpy out: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    del P, D

# the same for de(single=4) at the CRISPRi-screen shape (config 3)
del dt
p = synth.device_problem(1002, 10000, 50000, dev, n_group=300, group_p=0.02, n_module=0)
h = {k: p[k].cpu().numpy() for k in ("dg", "dt", "dc")}
del p
torch.cuda.empty_cache()
for rep in range(3):
    t0 = time.perf_counter()
    res = norm.de(h["dg"], h["dt"], h["dc"], single=4)
    print("de(single=4), numpy in / numpy out: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
