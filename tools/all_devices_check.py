#!/usr/bin/env python3
"""ONE process, every visible GPU: ``norm.coex(dt, dc, devices="all")`` / ``norm.de(..., devices="all")`` must
return the reference's complete return value, bit-identical to the single-GPU call.

    python tools/all_devices_check.py [genes] [cells] [--time]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import normalisr_b200.normalisr as norm  # noqa: E402
from normalisr_b200 import synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    genes = int(args[0]) if len(args) > 0 else 5000
    cells = int(args[1]) if len(args) > 1 else 3000
    n_dev = torch.cuda.device_count()
    p = synth.device_problem(77, genes, cells, torch.device("cuda", 0))
    dt = p["dt"].cpu().pin_memory()
    dc = p["dc"].cpu().numpy()
    del p
    torch.cuda.empty_cache()
    ok = True
    P1, D1, v1 = norm.coex(dt, dc)
    only_all = "--only-all" in sys.argv
    for devices in sorted(({n_dev} if only_all else {2, n_dev}) - {0, 1}):
        if devices > n_dev:
            continue
        for rep in range(2):
            P, D, v = norm.coex(dt, dc, devices=devices)
            good = bool(np.array_equal(P, P1) and np.array_equal(D, D1) and np.array_equal(v, v1))
            ok = ok and good
            print("all-devices check [coex, 1 process, %d GPUs, rep %d]: genes %d cells %d: complete (P, dot, var) "
                  "identical to single GPU: %s" % (devices, rep, genes, cells, good), flush=True)
    rng = np.random.default_rng(5)
    dg = (rng.random((37, cells)) < 0.05).astype(np.float64)
    dtn = dt.numpy()
    for single in (0, 4):
        r1 = norm.de(dg, dtn, dc, single=single)
        rN = norm.de(dg, dtn, dc, single=single, devices="all")
        good = all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(r1, rN))
        ok = ok and good
        print("all-devices check [de single=%d, 1 process, %d GPUs]: identical to single GPU: %s" % (single, n_dev, good),
              flush=True)
    if "--time" in sys.argv:
        out = (torch.empty((genes, genes), dtype=torch.float64, pin_memory=True),
               torch.empty((genes, genes), dtype=torch.float64, pin_memory=True))
        for devices in sorted({n_dev} if only_all else {1, 2, n_dev}):
            if devices > n_dev:
                continue
            ka = dict(devices=devices) if devices > 1 else {}
            for _ in range(2):
                norm.coex(dt, dc, out=out, **ka)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                norm.coex(dt, dc, out=out, **ka)
            sec = (time.perf_counter() - t0) / reps
            print("all-devices timing: norm.coex(host dt %dx%d, out=pinned) on %d GPU(s), one process: %.1f ms per call, "
                  "%.3e pairs/s (h2d %.2f GB, d2h %.2f GB per call)" % (
                      genes, cells, devices, 1e3 * sec, genes * (genes - 1) / 2 / sec, genes * cells * 8e-9,
                      genes * genes * 16e-9), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
