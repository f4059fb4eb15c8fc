#!/usr/bin/env python3
"""Golden fixtures for inv_rank with every option (reference src/normalisr/association.py:4-134: exact SVD,
rank cap ``mpc``, randomised truncated SVD with and without QR-normalised power iterations, stacks of
matrices) and for de(single=4) with an ``mpc`` that truncates nothing, made by the UNMODIFIED reference.

    python tests/golden/make_golden_inv_rank.py     # writes tests/golden/inv_rank.npz
"""
import json
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NSR_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
from normalisr.association import inv_rank  # noqa: E402  (the reference)
from normalisr.de import de  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OPTIONS = [{}, {"mpc": 5}, {"mpc": 5, "qr": 1}, {"mpc": 5, "qr": 3}, {"method": "scipy", "mpc": 4}, {"method": "sklearn"},
           {"method": "sklearn", "qr": 2}, {"mpc": 50}, {"tol": 1e-3}]


def main():
    rng = np.random.default_rng(77)
    out = {"options": json.dumps(OPTIONS)}
    for i, (n, r) in enumerate(((6, 6), (12, 7), (30, 30), (40, 11))):
        a = rng.normal(size=(n, r)) * rng.uniform(0.1, 10, size=(n, 1))
        m = a @ a.T
        out["m%d" % i] = m
        for j, ka in enumerate(OPTIONS):
            g, k = inv_rank(m, **ka)
            out["inv%d_%d" % (i, j)], out["rank%d_%d" % (i, j)] = g, np.int64(k)
    stack = np.stack([out["m0"], 3 * out["m0"]])
    g, k = inv_rank(stack)
    out["stack"], out["stack_inv"], out["stack_rank"] = stack, g, k
    # de(single=4): an mpc at least as large as the matrices inverted changes nothing
    n, ng, nt, nc = 300, 6, 20, 3
    dc = np.concatenate([rng.normal(size=(nc - 1, n)), np.ones((1, n))])
    dg = (rng.random((ng, n)) < 0.15).astype(float)
    dt = rng.normal(size=(nt, n)) + 0.4 * dg[2] + 0.3 * dc[0]
    r0 = de(dg, dt, dc, single=4)
    r1 = de(dg, dt, dc, single=4, mpc=ng - 1 + nc, method="scipy")
    assert all(np.array_equal(a, b) for a, b in zip(r0, r1) if a is not None)
    out.update(de_dg=dg, de_dt=dt, de_dc=dc, de_P=r1[0], de_gamma=r1[1], de_varg=r1[3], de_vart=r1[4])
    np.savez_compressed(os.path.join(HERE, "inv_rank.npz"), **out)
    print("inv_rank", len(OPTIONS), "option sets x 4 matrices; de(single=4, mpc) min P", float(r1[0].min()))


if __name__ == "__main__":
    main()
