#!/usr/bin/env python3
"""Golden fixtures for de(single=1) (low-MOI screens; reference association_test_2,
src/normalisr/association.py:263-390, driver :910-925), made by the UNMODIFIED reference.

    python tests/golden/make_golden_single1.py     # writes tests/golden/de_single1*.npz
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


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def low_moi(rng, n_grna, n_cell, p_none=0.45, p_double=0.08):
    """Low-MOI design: most cells carry no or one gRNA, a few carry two."""
    dg = np.zeros((n_grna, n_cell))
    u = rng.random(n_cell)
    one = (u >= p_none)
    dg[rng.integers(0, n_grna, size=n_cell)[one], np.nonzero(one)[0]] = 1
    two = np.nonzero(u >= 1 - p_double)[0]
    dg[rng.integers(0, n_grna, size=two.size), two] = 1
    return dg


def main():
    rng = np.random.default_rng(515)
    n, g, k = 1500, 60, 12
    dc = np.concatenate([rng.normal(size=(3, n)), (rng.random((1, n)) < 0.4).astype(float), np.ones((1, n))])
    dg = low_moi(rng, k, n)
    dg[5] = 0                                   # constant grouping: dropped and back-filled by de
    dt = rng.normal(size=(g, n)) + 0.4 * dc[0] + 2.0
    for x in range(k):                          # planted effects of different sizes -> P from ~1 down to 1e-60
        dt[x] += 0.35 * x * dg[x]
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc, single=1)
    save("de_single1", dg=dg, dt=dt, dc=dc, P=P, gamma=gamma, varg=varg, vart=vart)
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc, single=1, lowmem=False, dimreduce=1)
    save("de_single1_alpha", dg=dg, dt=dt, dc=dc, P=P, gamma=gamma, alpha=alpha, varg=varg, vart=vart,
         dimreduce=np.int64(1))
    # rank-deficient covariates within some subsets + no covariates at all
    # (a covariate equal to a tested grouping would leave rounding noise / rounding noise in the
    # reference, association.py:359-361 "should never happen in theory": not a usable fixture)
    dc2 = np.concatenate([dc, dc[:1] - 2 * dc[4:5], 1 - dc[3:4]])
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc2, single=1)
    save("de_single1_rankdef", dg=dg, dt=dt, dc=dc2, P=P, gamma=gamma, varg=varg, vart=vart)
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc[:0], single=1)
    save("de_single1_nocov", dg=dg, dt=dt, dc=dc[:0], P=P, gamma=gamma, varg=varg, vart=vart)


if __name__ == "__main__":
    main()
