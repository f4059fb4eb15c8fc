"""The oracle (oracle/normalisr_oracle.py) against every golden vector produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import normalisr_oracle as orc
from conftest import assert_p_close, load_golden, pearson_from, R_ATOL


@pytest.mark.parametrize("case", ["coex_chain", "coex_tail", "coex_rankdef", "coex_nocov"])
def test_coex_matches_reference(case):
    g = load_golden(case)
    ka = {"dimreduce": int(g["dimreduce"])} if "dimreduce" in g else {}
    P, dot, var = orc.coex(g["dt"], g["dc"], **ka)
    # the oracle follows the reference operation by operation: agreement is at rounding level
    np.testing.assert_allclose(var, g["var"], rtol=1e-12)
    np.testing.assert_allclose(dot, g["dot"], rtol=1e-9, atol=1e-13)
    assert_p_close(P, g["P"], rtol=1e-9)
    r = pearson_from(dot, var, var)
    assert np.abs(r - pearson_from(g["dot"], g["var"], g["var"])).max() <= R_ATOL * 1e-3
    assert (np.diag(P) == 0).all() and (np.diag(dot) == 0).all()


def test_coex_tile_size_invariance():
    g = load_golden("coex_chain")
    a = orc.coex(g["dt"], g["dc"])
    b = orc.coex(g["dt"], g["dc"], bsx=7)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-10)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-15)


def test_coex_threads_match_serial():
    g = load_golden("coex_chain")
    a = orc.coex(g["dt"], g["dc"], bsx=16)
    b = orc.coex(g["dt"], g["dc"], bsx=16, nth=3)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("case,ka", [("de_single0", {}), ("de_single0_alpha", {"lowmem": False}),
                                     ("de_single4", {"single": 4}),
                                     ("de_single4_rankdef", {"single": 4}),
                                     ("de_single1", {"single": 1}),
                                     ("de_single1_alpha", {"single": 1, "lowmem": False, "dimreduce": 1}),
                                     ("de_single1_rankdef", {"single": 1}),
                                     ("de_single1_nocov", {"single": 1})])
def test_de_matches_reference(case, ka):
    g = load_golden(case)
    P, gamma, alpha, varg, vart = orc.de(g["dg"], g["dt"], g["dc"], **ka)
    np.testing.assert_allclose(varg, g["varg"], rtol=1e-9)
    np.testing.assert_allclose(vart, g["vart"], rtol=1e-9)
    np.testing.assert_allclose(gamma, g["gamma"], rtol=1e-7, atol=1e-12)
    assert_p_close(P, g["P"], rtol=1e-7)
    if "alpha" in g:
        np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-7, atol=1e-10)
    else:
        assert alpha is None
    # the constant grouping (row 5) is back-filled: P=1, everything else 0 (de.py:107-122)
    assert (P[5] == 1).all() and (gamma[5] == 0).all() and varg[5] == 0 and (vart[5] == 0).all()


def test_pvalue_known_answers_scipy_and_c():
    g = load_golden("pvalue_kat")
    ps = orc.beta_cdf(1 - g["r2"], g["a"], backend="scipy")
    assert_p_close(ps, g["P"], rtol=1e-12)
    pc = orc.beta_cdf(1 - g["r2"], g["a"], backend="c")
    assert_p_close(pc, g["P"], rtol=1e-9)


def test_pvalue_mpmath_tail_points():
    """60-digit mpmath values of I_{1-r^2}((n-1-c)/2, 1/2)."""
    pts = [(2000, 5, 0.05, 0.025494348631248), (100000, 5, 0.117, 1.1098053295708e-301),
           (10000, 5, 0.3, 5.6798583929825e-207), (1000000, 5, 0.037, 7.1939093799116e-300)]
    for n, c, r, want in pts:
        a = (n - 1 - c) / 2
        for backend in ("scipy", "c"):
            got = float(np.ravel(orc.beta_cdf(1 - r * r, a, backend=backend))[0])
            assert abs(got - want) / want < 1e-9, (n, r, backend, got, want)


def test_pinv_rank_tolerance():
    rng = np.random.default_rng(0)
    c = rng.normal(size=(4, 50))
    c = np.concatenate([c, c[:1] * 2 - c[1:2]])
    mi, r = orc.pinv_rank(c @ c.T)
    assert r == 4
    np.testing.assert_allclose(mi, np.linalg.pinv(c @ c.T, rcond=1e-8), atol=1e-10)


def test_errors():
    with pytest.raises(ValueError):
        orc.coex(np.zeros((3, 4)), np.random.default_rng(0).normal(size=(3, 4)))  # n <= rank + 1
    with pytest.raises(ValueError):
        orc.association_tests(np.zeros((3, 9)), None, np.ones((1, 9)), single=7)


def test_oracle_single4_same_matches_reference():
    """association_tests(dx, None, dc, single=4): every pair with all other rows as covariates
    (association.py:492-496, 517-556, 1036-1065)."""
    g = load_golden("single4_same")
    iu = np.triu_indices(g["dx"].shape[0], 1)
    for name, ka in (("lowmem", {}), ("alpha", dict(lowmem=False)), ("gamma", dict(return_dot=False)),
                     ("dimreduce", dict(dimreduce=3))):
        r = orc.association_tests(g["dx"], None, g["dc"], single=4, **ka)
        np.testing.assert_allclose(r[0][iu], g["P_" + name][iu], rtol=1e-10)
        np.testing.assert_allclose(r[1], g["dot_" + name], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(r[4], g["vary_" + name], rtol=1e-10)
        assert r[3] is None and (r[0].diagonal() == 0).all()
        if not ka.get("lowmem", True):
            np.testing.assert_allclose(r[2], g["alpha_" + name], rtol=1e-9, atol=1e-12)
    r = orc.association_tests(g["dx"], None, g["dc2"], single=4, lowmem=False)
    np.testing.assert_allclose(r[0][iu], g["P_rankdef"][iu], rtol=1e-10)
    np.testing.assert_allclose(r[2], g["alpha_rankdef"], rtol=1e-9, atol=1e-12)
