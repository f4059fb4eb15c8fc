#!/usr/bin/env python3
"""Golden fixture for compute_var (reference src/normalisr/norm.py:56-128), made by
the UNMODIFIED reference on the output of its own lcpm / normcov.

    python tests/golden/make_golden_compute_var.py     # writes tests/golden/compute_var.npz
"""
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NSR_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
import normalisr.normalisr as norm  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import nb_counts, batches  # noqa: E402


def main():
    rng = np.random.default_rng(606)
    reads = nb_counts(rng, 200, 400, 20)
    dt, _, _, dcov = norm.lcpm(reads, nth=1)
    dc = norm.normcov(np.concatenate([batches(rng, reads.shape[1], 3), dcov], axis=0))
    dt, dc = np.ascontiguousarray(dt), np.ascontiguousarray(dc)
    w = norm.compute_var(dt, dc)
    dc2 = np.concatenate([dc, dc[:1] + 2 * dc[-1:]])               # rank-deficient covariates
    w2 = norm.compute_var(dt, dc2)
    w3 = norm.compute_var(dt, dc, stepmax=4)                       # EM-like iterations (norm.py:97-121)
    w4 = norm.compute_var(dt, dc, stepmax=50, eps=1e-3)            # stops on the tolerance
    np.savez_compressed(os.path.join(HERE, "compute_var.npz"), dt=dt, dc=dc, w=w, dc2=dc2, w2=w2, w_step4=w3, w_eps=w4)
    print("compute_var", dt.shape, dc.shape, w[:4])


if __name__ == "__main__":
    main()
