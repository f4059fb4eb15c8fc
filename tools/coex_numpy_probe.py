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

# the steps around the hot path on plain numpy arrays (50k cells x 10k genes)
del h
rng = np.random.default_rng(3)
n, g = 50000, 10000
counts = rng.poisson(2.0, size=(g, n)).astype(np.int64)
for rep in range(2):
    t0 = time.perf_counter()
    lc, _, _, cov = norm.lcpm(counts)
    print("lcpm, numpy in / numpy out: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
cov = np.concatenate([cov, np.ones((1, n))])
for rep in range(2):
    t0 = time.perf_counter()
    w = norm.compute_var(lc, cov)
    t1 = time.perf_counter()
    out = norm.normvar(lc, cov, w, np.full(g, 0.5))
    print("compute_var %.1f ms, normvar %.1f ms (numpy in / numpy out)" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), flush=True)
P = np.random.default_rng(4).random((20000, 20000)) ** 3
for rep in range(2):
    t0 = time.perf_counter()
    net = norm.binnet(P, 0.05)
    print("binnet 20k x 20k, numpy in / numpy out: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
