#!/usr/bin/env python3
"""How many digit products does the stated tolerance need?  A numpy model of the fixed-point contraction (runs on
the CPU): projection in float64, the sign-randomised 128-point Walsh-Hadamard mix, quantisation of every row to
signed digits (quantum = 6 rms / full scale, as csrc/residual.cu), exact digit-plane products (float64 dgemm on
small integers is exact), the kept products combined with their weights - against the float64 oracle, on the
tail problems of tests/test_gpu_parity.py (P down to 1e-300).  Schemes:

  3 x 8 bits, all 9 / 8 (default) / 7 (drop (2,1) only: the asymmetric variant) / 6 ("fast") products
  3 x 7 bits, all 9 products (what a Karatsuba arrangement computes with 6 tensor-core products)
  4 x 8 bits, 10 products ("precise")

    python tools/precision_probe.py [cells genes]      # default 100000 300
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import normalisr_oracle as orc          # noqa: E402
import nsr_testlib as tl                # noqa: E402
from normalisr_b200.association import covariate_basis       # noqa: E402  (host numpy function)


def digit_planes(v, n_digits, bits):
    """balanced base-2^bits digits of integer-valued v, most significant first"""
    base, half = 1 << bits, 1 << (bits - 1)
    out, r = [], v.copy()
    for _ in range(n_digits):
        d = ((r + half) % base) - half
        out.append(d)
        r = (r - d) / base
    assert np.abs(r).max() == 0, "value beyond the digit range"
    return out[::-1]


def scheme(zp, n_digits, bits, keep):
    """sum_k V_i V_j from the kept digit-pair products, scaled back to the units of z'."""
    full = sum((1 << (bits - 1)) - 1 if d == 0 else 0 for d in range(1))       # top digit magnitude
    vmax = 0
    for d in range(n_digits):
        vmax = vmax * (1 << bits) + ((1 << (bits - 1)) - 1)
    rms = np.sqrt((zp * zp).mean(axis=1))
    quantum = 6.0 * rms / vmax
    v = np.rint(zp / quantum[:, None])
    over = np.abs(v).max(axis=1) > vmax
    if over.any():                                   # the fix-up pass: quantum = max / full scale for those rows
        quantum[over] = np.abs(zp[over]).max(axis=1) / vmax
        v[over] = np.rint(zp[over] / quantum[over, None])
    planes = digit_planes(v, n_digits, bits)
    tot = np.zeros((zp.shape[0], zp.shape[0]))
    for a in range(n_digits):
        for b in range(n_digits):
            if keep(a, b):
                w = float(1 << (bits * (2 * n_digits - 2 - a - b)))
                tot += w * (planes[a] @ planes[b].T)      # exact: |digit products| <= 2^14, sums < 2^53
    return tot * np.outer(quantum, quantum)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rng = np.random.default_rng(4100)
    b = rng.integers(0, 6, size=n)
    dc = np.array([(b == i).astype(float) for i in range(1, 6)] + [rng.normal(size=n) for _ in range(3)] + [np.ones(n)])
    hi = float(np.sqrt(1380.0 / n)) * 1.25                 # loadings that put the strongest pairs just past P = 1e-300
    a = rng.uniform(0.35 * hi ** 0.5, hi ** 0.5 * 1.05, size=g)
    a[:g // 5] = 0.0
    dt = rng.normal(size=(g, n)) + a[:, None] * rng.normal(size=n)[None, :]
    dt += 0.3 * dc[5][None, :] + rng.uniform(2, 6, size=(g, 1))
    P_ref, dot_ref, var_ref = orc.coex(dt, dc)
    Qt, rank, _ = covariate_basis(dc)
    z = tl.residual(dt, Qt)
    var = (z * z).mean(axis=1)
    zp = tl.hadamard128(z)
    iu = np.triu_indices(g, 1)
    r_ref = (dot_ref / np.sqrt(np.outer(var_ref, var_ref)))[iu]
    band = (P_ref[iu] >= 1e-300) & (P_ref[iu] <= 1e-200)
    print("cells %d, genes %d: %d pairs, %d with 1e-300 <= P <= 1e-200, max |r| %.4f" % (n, g, r_ref.size, band.sum(), np.abs(r_ref).max()))
    print("%-44s %12s %12s %14s" % ("scheme", "max |dr|", "rms dr", "max rel dP"))
    dof = (n - 1 - rank) / 2
    schemes = [
        ("3 x 8 bits, 9 products", 3, 8, lambda a_, b_: True),
        ("3 x 8 bits, 8 products (default)", 3, 8, lambda a_, b_: a_ + b_ <= 3),
        ("3 x 8 bits, 7 products (without (2,1))", 3, 8, lambda a_, b_: a_ + b_ <= 3 and (a_, b_) != (2, 1)),
        ("3 x 8 bits, 6 products (fast)", 3, 8, lambda a_, b_: a_ + b_ <= 2),
        ("3 x 7 bits, 9 products (Karatsuba: 6 MMAs)", 3, 7, lambda a_, b_: True),
        ("4 x 8 bits, 10 products (precise)", 4, 8, lambda a_, b_: a_ + b_ <= 3),
    ]
    for name, nd, bits, keep in schemes:
        s = scheme(zp, nd, bits, keep)
        dot = s / n
        r = (dot / np.sqrt(np.outer(var, var)))[iu]
        r2 = np.minimum(r * r, 1.0)
        P = orc.beta_cdf(1.0 - r2, dof)
        ok = P_ref[iu] >= 1e-300
        rel = np.abs(P[ok] - P_ref[iu][ok]) / P_ref[iu][ok]
        print("%-44s %12.3e %12.3e %14.3e" % (name, np.abs(r - r_ref).max(), np.sqrt(((r - r_ref) ** 2).mean()), rel.max()))


if __name__ == "__main__":
    main()
