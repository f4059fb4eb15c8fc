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
