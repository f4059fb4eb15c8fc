#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (lingfeiwang/normalisr v1.0.0, imported from /root/reference/src) in the
authoring container.  /root/reference does not exist on the GPU box, so the
vectors are committed; this script is what made them.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Every case stores the exact inputs handed to normalisr.normalisr.coex / de and the
arrays they returned (float64), so both the oracle (oracle/normalisr_oracle.py)
and the CUDA path can be checked against the reference on identical inputs.
Library versions at generation time are recorded in tests/golden/MANIFEST.json.
"""
import json
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NSR_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
import normalisr.normalisr as norm  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))


def nb_counts(rng, n_gene, n_cell, min_cells):
    """Synthetic NB count matrix, SURVEY.md section 8(d) recipe."""
    mu = rng.gamma(0.5, 2.0, size=n_gene) + 0.05
    depth = rng.lognormal(0.0, 0.4, size=n_cell)
    m = mu[:, None] * depth[None, :]
    reads = rng.negative_binomial(2, 2.0 / (2.0 + m))
    keep = (reads > 0).sum(axis=1) >= min_cells
    return reads[keep].astype("u8")


def batches(rng, n_cell, n_batch):
    b = rng.integers(0, n_batch, size=n_cell)
    return np.array([(b == i) for i in range(1, n_batch)], dtype=float)


def chain(reads, cov_cat):
    """reference lcpm -> normcov -> scaling_factor -> compute_var -> normvar
    (examples/GSE123139/code/cmd_coex.sh steps 3-8, without the QC steps)."""
    dt, _, _, dcov = norm.lcpm(reads, nth=1)
    dc = np.concatenate([cov_cat, dcov], axis=0)
    dc = norm.normcov(dc)
    sf = norm.scaling_factor(reads)
    w = norm.compute_var(dt, dc)
    dtn, dcn = norm.normvar(dt, dc, w, sf, nth=1)
    return np.ascontiguousarray(dtn), np.ascontiguousarray(dcn)


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def main():
    manifest = {}
    # ---- case 1: full chain, coex ------------------------------------------------
    rng = np.random.default_rng(1001)
    reads = nb_counts(rng, 70, 400, 20)
    dt, dc = chain(reads, batches(rng, 400, 4))
    P, dot, var = norm.coex(dt, dc)
    save("coex_chain", dt=dt, dc=dc, P=P, dot=dot, var=var)

    # ---- case 2: coex, planted modules so the P tail reaches ~1e-300 ---------------
    rng = np.random.default_rng(1002)
    n, g = 2000, 40
    dc = np.concatenate([rng.normal(size=(3, n)), np.ones((1, n))])
    dt = rng.normal(size=(g, n)) + 5.0
    f = rng.normal(size=n)
    for i, s in enumerate(np.linspace(0.05, 1.6, 24)):
        dt[i] += s * f
    dt[30] = dt[31] * (1 + 1e-9)          # |r| ~ 1: P underflows to 0 in the reference
    dt += 0.3 * dc[0]
    P, dot, var = norm.coex(dt, dc)
    save("coex_tail", dt=dt, dc=dc, P=P, dot=dot, var=var)

    # ---- case 3: rank-deficient covariates + dimreduce kwarg ---------------------
    rng = np.random.default_rng(1003)
    n, g = 500, 33
    base = rng.normal(size=(3, n))
    dc = np.concatenate([base, base[:1] + base[1:2], np.ones((1, n))])   # rank 4 of 5
    dt = rng.normal(size=(g, n)) * rng.uniform(0.5, 3, size=(g, 1)) + rng.normal(size=(g, 1))
    dt[7] = 0.0                                                          # var 0 -> 1 rule (:231)
    P, dot, var = norm.coex(dt, dc, dimreduce=2)
    save("coex_rankdef", dt=dt, dc=dc, P=P, dot=dot, var=var, dimreduce=np.int64(2))

    # ---- case 4: no covariates at all ---------------------------------------------
    rng = np.random.default_rng(1004)
    n, g = 257, 19
    dt = rng.normal(size=(g, n))
    dc = np.zeros((0, n))
    P, dot, var = norm.coex(dt, dc)
    save("coex_nocov", dt=dt, dc=dc, P=P, dot=dot, var=var)

    # ---- case 5: de single=0 (one constant grouping is dropped and back-filled) ----
    rng = np.random.default_rng(1005)
    reads = nb_counts(rng, 60, 600, 30)
    dt, dc = chain(reads, batches(rng, 600, 3))
    dg = (rng.random(size=(12, 600)) < 0.08).astype(float)
    dg[5] = 0.0
    for i in range(4):                                   # true effects
        dt[i] += 0.8 * dg[i]
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc)
    assert alpha is None
    save("de_single0", dg=dg, dt=dt, dc=dc, P=P, gamma=gamma, varg=varg, vart=vart)
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc, lowmem=False)
    save("de_single0_alpha", dg=dg, dt=dt, dc=dc, P=P, gamma=gamma, alpha=alpha,
         varg=varg, vart=vart)

    # ---- case 6: de single=4 (other gRNAs as covariates) ---------------------------
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc, single=4)
    save("de_single4", dg=dg, dt=dt, dc=dc, P=P, gamma=gamma, varg=varg, vart=vart)
    # rank-deficient covariates: the per-x pseudo-inverse drops one direction, so the
    # rank that enters the degrees of freedom is m-2, not m-1.  (Exactly collinear
    # gRNA rows make the reference itself fail its own assert at association.py:557,
    # so that case has no golden vector.)
    dc2 = np.concatenate([dc[:2], dc[:1] - 2 * dc[1:2], dc[2:]])
    P, gamma, alpha, varg, vart = norm.de(dg, dt, dc2, single=4)
    save("de_single4_rankdef", dg=dg, dt=dt, dc=dc2, P=P, gamma=gamma, varg=varg,
         vart=vart)

    # ---- case 7: P-value known answers (scipy.stats.beta.cdf as the reference calls it)
    from scipy.stats import beta
    rng = np.random.default_rng(1007)
    a = np.array([0.5, 1.0, 2.5, 7.0, 14.5, 15.0, 40.0, 123.5, 995.5, 4995.0, 24995.5,
                  49995.0, 499995.0])
    r2 = np.concatenate([[0.0, 1e-300, 1e-18, 1e-12, 1e-9, 1e-6, 1e-4, 1e-3, 0.01, 0.05, 0.1,
                          0.2, 0.29, 0.3, 0.31, 0.5, 0.7, 0.9, 0.99, 0.999999, 1.0],
                         10 ** rng.uniform(-8, 0, size=60)])
    A, R2 = np.meshgrid(a, r2, indexing="ij")
    pv = beta.cdf(1 - R2, A, 0.5)
    save("pvalue_kat", a=A, r2=R2, P=pv)

    import scipy
    import sklearn
    manifest = {
        "reference": "lingfeiwang/normalisr v1.0.0 (setup.py:6) imported from " + REF,
        "numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__,
        "python": sys.version.split()[0],
    }
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
