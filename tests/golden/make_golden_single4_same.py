#!/usr/bin/env python3
"""Golden fixtures for association_tests(dx, None, dc, single=4) - every pair of rows tested with all
other rows as covariates (reference src/normalisr/association.py:492-496, 517-556, 1036-1065) - made by
the UNMODIFIED reference.

    python tests/golden/make_golden_single4_same.py     # writes tests/golden/single4_same*.npz
"""
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NSR_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
from normalisr.association import association_tests  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(909)
    n, nx, nc = 400, 14, 4
    dc = np.concatenate([rng.normal(size=(nc - 1, n)), np.ones((1, n))])
    f = rng.normal(size=(3, n))
    dx = rng.normal(size=(nx, n)) + rng.normal(size=(nx, 3)) @ f + 0.5 * rng.normal(size=(nx, nc)) @ dc
    out = {"dx": dx, "dc": dc}
    for name, ka in (("lowmem", dict()), ("alpha", dict(lowmem=False)), ("gamma", dict(return_dot=False)),
                     ("dimreduce", dict(dimreduce=3))):
        r = association_tests(dx, None, dc, single=4, **ka)
        assert r[3] is None
        out["P_" + name], out["dot_" + name], out["vary_" + name] = r[0], r[1], r[4]
        if r[2] is not None:
            out["alpha_" + name] = r[2]
    dc2 = np.concatenate([dc, dc[:1] - 2 * dc[1:2]])                # rank-deficient covariates
    r = association_tests(dx, None, dc2, single=4, lowmem=False)
    out.update(dc2=dc2, P_rankdef=r[0], dot_rankdef=r[1], alpha_rankdef=r[2], vary_rankdef=r[4])
    np.savez_compressed(os.path.join(HERE, "single4_same.npz"), **out)
    print("single4_same", dx.shape, float(out["P_lowmem"][np.triu_indices(nx, 1)].min()))


if __name__ == "__main__":
    main()
