#!/usr/bin/env python3
"""Golden fixtures for lcpm (reference src/normalisr/lcpm.py:21-208), made by the
UNMODIFIED reference.

    python tests/golden/make_golden_lcpm.py     # writes tests/golden/lcpm_*.npz
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
from make_golden import nb_counts  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    reads = nb_counts(rng, 260, 300, 10)
    reads[3, 5] = 70000                              # one large count: long look-up table
    dt, dmean, dvar, dcov = norm.lcpm(reads, nth=1)
    dt2, dmean2, dvar2, _ = norm.lcpm(reads, nth=1, lowmem=False, nocov=True)
    dt3, _, _, dcov3 = norm.lcpm(reads, nth=1, normalize=False, ntot=10**9)
    np.savez_compressed(os.path.join(HERE, "lcpm_counts.npz"), reads=reads, lcpm=dt, cov=dcov, mean=dmean2, var=dvar2,
                        lcpm_raw=dt3, cov_raw=dcov3)
    print("lcpm_counts", reads.shape, reads.dtype, dt.shape)
    # posterior resampling (varscale != 0, lcpm.py:104-109, 134-150, 178-190): the reference seeds numpy's
    # global generator (:82-83) and draws one randn(n_gene, n_cell)
    rng = np.random.default_rng(2025)
    reads = nb_counts(rng, 120, 150, 10)
    varscale, seed = 0.6, 12345
    r1 = norm.lcpm(reads, nth=1, varscale=varscale, seed=seed)
    r2 = norm.lcpm(reads, nth=1, varscale=varscale, seed=seed, lowmem=False)
    np.savez_compressed(os.path.join(HERE, "lcpm_resample.npz"), reads=reads, varscale=varscale, seed=seed,
                        lcpm=r1[0], cov=r1[3], lcpm_full=r2[0], mean_full=r2[1], var_full=r2[2])
    print("lcpm_resample", reads.shape, float(np.abs(r1[0] - r2[0]).max()))


if __name__ == "__main__":
    main()
