#!/usr/bin/env python3
"""The steps around the hot path on one GPU, each timed on device-resident inputs (the blocks of bench.py's
line): lcpm, compute_var, normvar, binnet.  Also the driver for `ncu -k regex:...` captures of their kernels.

    python tools/aux_run.py [lcpm] [compute_var] [normvar] [binnet]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from normalisr_b200 import association, engine, synth  # noqa: E402


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["lcpm", "compute_var", "normvar", "binnet"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}
    if "lcpm" in which:
        out["lcpm"] = bench.bench_lcpm(torch, dev)
    if "compute_var" in which:
        out["compute_var"] = bench.bench_compute_var(torch, dev)
    if "normvar" in which:
        out["normvar"] = bench.bench_normvar(torch, dev)
    if "binnet" in which:
        ctx = engine.context(0)
        n_gene, n_cell = 20000, 20000
        p = synth.device_problem(1004, n_gene, n_cell, dev)
        res = association.association_tests(p["dt"], None, p["dc"])
        P = res[0]
        del p
        out["binnet"] = bench.bench_binnet(torch, ctx, P, n_gene)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
