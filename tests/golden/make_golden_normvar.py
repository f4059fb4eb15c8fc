#!/usr/bin/env python3
"""Golden fixtures for normvar (reference src/normalisr/norm.py:131-289), made by
the UNMODIFIED reference on the output of its own lcpm -> normcov -> scaling_factor ->
compute_var chain.

    python tests/golden/make_golden_normvar.py     # writes tests/golden/normvar_*.npz
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


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def main():
    rng = np.random.default_rng(909)
    reads = nb_counts(rng, 150, 500, 20)[:90]
    cov_cat = batches(rng, reads.shape[1], 3)
    dt, _, _, dcov = norm.lcpm(reads, nth=1)
    dc = norm.normcov(np.concatenate([cov_cat, dcov], axis=0))
    sf = norm.scaling_factor(reads)
    w = norm.compute_var(dt, dc)
    dt = np.ascontiguousarray(dt)
    dc = np.ascontiguousarray(dc)
    sf[7] = 0.0                                   # a gene that is not rescaled (norm.py:239)
    dtn, dcn = norm.normvar(dt, dc, w, sf, nth=1)
    save("normvar_chain", dt=dt, dc=dc, w=w, wt=sf, dtn=dtn, dcn=dcn)
    dextra = rng.normal(size=(2, dt.shape[1]))
    for cat in (0, 2):
        dtn, dcn, dxn = norm.normvar(dt, dc, w, sf, dextra=dextra, cat=cat, keepvar=False, normmean=True, nth=1)
        save("normvar_cat%d" % cat, dt=dt, dc=dc, w=w, wt=sf, dextra=dextra, dtn=dtn, dcn=dcn, dextran=dxn,
             cat=np.int64(cat))


if __name__ == "__main__":
    main()
