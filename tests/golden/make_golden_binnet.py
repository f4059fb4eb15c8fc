#!/usr/bin/env python3
"""Golden fixtures for the binnet row: run the UNMODIFIED reference
(normalisr.binnet.bh / binnet, src/normalisr/binnet.py) in the authoring container.

    python tests/golden/make_golden_binnet.py      # writes tests/golden/binnet_*.npz
"""
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NSR_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
import normalisr.normalisr as norm  # noqa: E402  (the reference)
from normalisr import binnet as ref_binnet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
QCUTS = np.array([0.5, 0.05, 1e-3, 1e-12])


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def nets(P):
    out = []
    for q in QCUTS:
        try:
            out.append(np.packbits(ref_binnet.binnet(P, q), axis=1))
        except RuntimeError:                      # "Empty binary network."
            out.append(np.zeros((P.shape[0], (P.shape[1] + 7) // 8), dtype=np.uint8))
    return np.array(out)


def main():
    rng = np.random.default_rng(4242)
    # 1. P-values of a real reference coex run: planted modules + a master regulator row
    n, g = 400, 300
    dc = np.concatenate([rng.normal(size=(2, n)), np.ones((1, n))])
    dt = rng.normal(size=(g, n))
    f = rng.normal(size=(3, n))
    dt[:40] += 0.9 * f[0]
    dt[40:70] += 0.5 * f[1]
    dt[70:150] += 0.2 * f[2]
    dt[299] = 0.6 * f[0] + 0.6 * f[1] + 0.4 * rng.normal(size=n)
    P, _, _ = norm.coex(dt, dc)
    save("binnet_coex", P=P, qcut=QCUTS, net=nets(P))
    # 2. ties, exact zeros and ones, constant rows (symmetry is not required by binnet)
    vals = np.array([0.0, 1e-300, 1e-9, 1e-4, 0.003, 0.003, 0.01, 0.2, 0.5, 1.0])
    T = vals[rng.integers(0, vals.size, size=(130, 130))]
    T[5] = 0.25
    T[6] = 0.0
    T[7] = 1.0
    T[8, :60] = 1e-5
    save("binnet_ties", P=T, qcut=QCUTS, net=nets(T))
    # 3. bh known answers (with and without ties)
    pv = [rng.random(257), vals[rng.integers(0, vals.size, size=300)], np.array([0.5]),
          np.sort(rng.random(64)) ** 3, np.zeros(9), np.ones(5)]
    save("bh_kat", **{"p%d" % i: x for i, x in enumerate(pv)},
         **{"q%d" % i: ref_binnet.bh(x.copy()) for i, x in enumerate(pv)}, count=np.int64(len(pv)))


if __name__ == "__main__":
    main()
