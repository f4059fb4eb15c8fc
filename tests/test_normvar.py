"""normvar (reference src/normalisr/norm.py:131-289).  CPU: the oracle against the
golden vectors made by the unmodified reference.  GPU: the CUDA path against both."""
import numpy as np
import pytest
import torch

import normalisr_oracle as orc
from conftest import load_golden

gpu = pytest.mark.gpu


def _args(g):
    ka = {}
    if "dextra" in g:
        ka = dict(dextra=g["dextra"], cat=int(g["cat"]), keepvar=False, normmean=True)
    return (g["dt"], g["dc"], g["w"], g["wt"]), ka


@pytest.mark.parametrize("case", ["normvar_chain", "normvar_cat0", "normvar_cat2"])
def test_oracle_normvar_matches_reference(case):
    g = load_golden(case)
    a, ka = _args(g)
    out = orc.normvar(*a, **ka)
    np.testing.assert_allclose(out[0], g["dtn"], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(out[1], g["dcn"])
    if "dextran" in g:
        np.testing.assert_array_equal(out[2], g["dextran"])
    with pytest.raises(ValueError):
        orc.normvar(g["dt"], g["dc"][:0], g["w"], g["wt"])
    with pytest.raises(ValueError):
        orc.normvar(g["dt"], g["dc"], -g["w"], g["wt"])


@gpu
@pytest.mark.parametrize("case", ["normvar_chain", "normvar_cat0", "normvar_cat2"])
def test_normvar_golden(case):
    from normalisr_b200 import normalisr as norm
    g = load_golden(case)
    a, ka = _args(g)
    out = norm.normvar(*a, **ka)
    assert all(isinstance(x, np.ndarray) and x.dtype == np.float64 for x in out)
    np.testing.assert_allclose(out[0], g["dtn"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out[1], g["dcn"], rtol=1e-15)
    if "dextran" in g:
        np.testing.assert_allclose(out[2], g["dextran"], rtol=1e-15)
    dev = norm.normvar(*[torch.from_numpy(x).cuda() for x in a],
                       **{k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in ka.items()})
    assert all(x.is_cuda for x in dev)
    np.testing.assert_allclose(dev[0].cpu().numpy(), g["dtn"], rtol=1e-9, atol=1e-9)


@gpu
@pytest.mark.parametrize("chebyshev", [True, False])
def test_normvar_against_oracle_shapes_and_chunks(monkeypatch, chebyshev):
    """Both sources of the per-gene Gram matrices: interpolation in wt (default; also the only one beyond 12
    covariates) and the streaming statistics pass."""
    from normalisr_b200 import norm as nv
    monkeypatch.setattr(nv, "_USE_CHEBYSHEV", chebyshev)
    rng = np.random.default_rng(41)
    for genes, n, nc in ((37, 300, 1), (500, 3000, 9), (260, 1001, 12), (64, 128, 5)) + (((90, 700, 16),) if chebyshev else ()):
        dc = np.concatenate([rng.normal(size=(nc - 1, n)), np.ones((1, n))]) if nc > 1 else np.ones((1, n))
        if nc >= 5:
            dc[1] = rng.random(n) < 0.3                       # a categorical covariate (cat = 1 leaves it alone)
            dc[2] = 2 * dc[0] - dc[-1]                         # rank-deficient covariates
        dt = rng.normal(size=(genes, n)) * rng.uniform(0.2, 3, size=(genes, 1)) - 6.0
        w = np.exp(rng.normal(size=n) * 0.3)
        wt = rng.uniform(0, 1.5, size=genes)
        wt[::7] = 0
        for ka in ({}, {"keepvar": False, "cat": 0}, {"normmean": True, "cat": 2, "dextra": rng.normal(size=(3, n))}):
            want = orc.normvar(dt, dc, w, wt, **ka)
            monkeypatch.setattr(nv, "_ROW_CHUNK_BYTES", 8 * n * 100)            # several row blocks
            got = nv.normvar(dt, dc, w, wt, **ka)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(b).max()))


def test_oracle_compute_var_matches_reference():
    g = load_golden("compute_var")
    np.testing.assert_allclose(orc.compute_var(g["dt"], g["dc"]), g["w"], rtol=1e-12)
    np.testing.assert_allclose(orc.compute_var(g["dt"], g["dc2"]), g["w2"], rtol=1e-12)
    np.testing.assert_allclose(orc.compute_var(g["dt"], g["dc"], stepmax=4), g["w_step4"], rtol=1e-12)
    np.testing.assert_allclose(orc.compute_var(g["dt"], g["dc"], stepmax=50, eps=1e-3), g["w_eps"], rtol=1e-12)
    assert np.abs(g["w_step4"] / g["w"] - 1).max() > 1e-3           # the iterations do move the result
    with pytest.raises(ValueError):
        orc.compute_var(g["dt"], g["dc"][:, :-1])


@gpu
def test_compute_var_golden_and_oracle(monkeypatch):
    from normalisr_b200 import norm as nv, normalisr as norm
    g = load_golden("compute_var")
    for dc, w in ((g["dc"], g["w"]), (g["dc2"], g["w2"])):
        got = norm.compute_var(g["dt"], dc)
        assert isinstance(got, np.ndarray) and got.shape == w.shape and got.min() == 1.0
        np.testing.assert_allclose(got, w, rtol=1e-10)
    dev = norm.compute_var(torch.from_numpy(g["dt"]).cuda(), torch.from_numpy(g["dc"]).cuda())
    assert dev.is_cuda
    np.testing.assert_allclose(dev.cpu().numpy(), g["w"], rtol=1e-10)
    # EM-like iterations (stepmax > 1, norm.py:97-121), fixed count and tolerance stop
    np.testing.assert_allclose(norm.compute_var(g["dt"], g["dc"], stepmax=4), g["w_step4"], rtol=1e-9)
    np.testing.assert_allclose(norm.compute_var(g["dt"], g["dc"], stepmax=50, eps=1e-3), g["w_eps"], rtol=1e-9)
    dev = norm.compute_var(torch.from_numpy(g["dt"]).cuda(), torch.from_numpy(g["dc"]).cuda(), stepmax=4)
    np.testing.assert_allclose(dev.cpu().numpy(), g["w_step4"], rtol=1e-9)
    rng = np.random.default_rng(14)
    n, genes = 3000, 900
    dc = np.concatenate([rng.normal(size=(4, n)), (rng.random((2, n)) < 0.3).astype(float), np.ones((1, n))])
    dt = rng.normal(size=(genes, n)) * np.exp(0.3 * dc[0]) * rng.uniform(0.5, 2, size=(genes, 1)) + 0.5 * dc[1] - 7.0
    want = orc.compute_var(dt, dc)
    monkeypatch.setattr(nv, "_ROW_CHUNK_BYTES", 8 * n * 200)                  # several row blocks
    np.testing.assert_allclose(norm.compute_var(dt, dc), want, rtol=1e-10)
    assert want.max() / want.min() > 1.5                                       # the cell-level trend is there
    with pytest.raises(ValueError):
        norm.compute_var(dt, dc[:, :-1])
    with pytest.raises(ValueError):
        norm.compute_var(dt, dc, eps=0)
    np.testing.assert_allclose(norm.compute_var(dt, dc, stepmax=3), orc.compute_var(dt, dc, stepmax=3), rtol=1e-9)   # chunked + iterated


@gpu
@pytest.mark.parametrize("chebyshev", [True, False])
def test_normvar_and_compute_var_many_covariates(monkeypatch, chebyshev):
    """More covariates than the kernels stage (16): same results through the library-product fallbacks
    (normvar: 20 and 33 covariates, rank-deficient; compute_var: covariate rank 24)."""
    from normalisr_b200 import norm as nv
    monkeypatch.setattr(nv, "_USE_CHEBYSHEV", chebyshev)
    monkeypatch.setattr(nv, "_WIDE_BYTES", 8 * 1200 * 64)                       # several gene chunks
    rng = np.random.default_rng(47)
    for genes, n, nc in ((150, 1200, 20), (70, 900, 33)):
        dc = np.concatenate([rng.normal(size=(nc - 1, n)), np.ones((1, n))])
        dc[1] = rng.random(n) < 0.3
        dc[2] = 2 * dc[0] - dc[-1]                                             # rank-deficient covariates
        dt = rng.normal(size=(genes, n)) * rng.uniform(0.2, 3, size=(genes, 1)) - 6.0
        w = np.exp(rng.normal(size=n) * 0.3)
        wt = rng.uniform(0, 1.5, size=genes)
        wt[::7] = 0
        for ka in ({}, {"keepvar": False, "cat": 0}, {"normmean": True, "cat": 2, "dextra": rng.normal(size=(3, n))}):
            want = orc.normvar(dt, dc, w, wt, **ka)
            monkeypatch.setattr(nv, "_ROW_CHUNK_BYTES", 8 * n * 100)
            got = nv.normvar(dt, dc, w, wt, **ka)
            for a, b in zip(got, want):
                np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(b).max()))
        dev = nv.normvar(*[torch.from_numpy(x).cuda() for x in (dt, dc, w, wt)])
        np.testing.assert_allclose(dev[0].cpu().numpy(), orc.normvar(dt, dc, w, wt)[0], rtol=1e-9, atol=1e-8)
    if chebyshev:
        n, genes = 2000, 500
        dc = np.concatenate([rng.normal(size=(23, n)), np.ones((1, n))])
        dt = rng.normal(size=(genes, n)) * np.exp(0.3 * dc[0]) * rng.uniform(0.5, 2, size=(genes, 1)) + 0.5 * dc[1] - 7.0
        monkeypatch.setattr(nv, "_ROW_CHUNK_BYTES", 8 * n * 200)
        for stepmax in (1, 3):
            np.testing.assert_allclose(nv.compute_var(dt, dc, stepmax=stepmax), orc.compute_var(dt, dc, stepmax=stepmax), rtol=1e-9)


@gpu
def test_normvar_errors():
    from normalisr_b200 import normalisr as norm
    dt, dc, w, wt = np.zeros((4, 10)), np.ones((2, 10)), np.ones(10), np.ones(4)
    for bad in ((dt, dc[:0], w, wt), (dt, dc[:, :9], w, wt), (dt, dc, w[:9], wt), (dt, dc, w, wt[:3]),
                (dt, dc, -w, wt), (dt, dc, w, -wt), (dt[0], dc, w, wt)):
        with pytest.raises(ValueError):
            norm.normvar(*bad)
    with pytest.raises(ValueError):
        norm.normvar(dt, dc, w, wt, cat=3)
    with pytest.raises(RuntimeError):
        norm.normvar(dt + 1, np.zeros((2, 10)), w, wt)        # zero-rank covariates (norm.py:161)


@gpu
def test_normvar_wide_weight_range_uses_more_pieces_or_falls_back():
    """Weights spanning orders of magnitude: the interpolation cuts the wt range into more pieces, and beyond
    its limit the Gram matrices come from the streaming pass - same results either way."""
    from normalisr_b200 import norm as nv
    rng = np.random.default_rng(43)
    genes, n, nc = 200, 1500, 6
    dc = np.concatenate([rng.normal(size=(nc - 1, n)), np.ones((1, n))])
    dt = rng.normal(size=(genes, n)) - 5.0
    for sd, top, fallback in ((1.5, 3.0, False), (3.0, 60.0, True)):
        w = np.exp(rng.normal(size=n) * sd)
        wt = rng.uniform(0, top, size=genes)
        G = nv._gram_chebyshev(torch.from_numpy(dc).cuda(), torch.log(torch.from_numpy(w).cuda()), torch.from_numpy(wt).cuda())
        assert (G is None) == fallback
        if not fallback:
            G, ok = G
            assert bool(ok)
            want = np.einsum('gk,ik,jk->gij', w[None, :] ** (2 * wt[:, None]), dc, dc)
            scale = np.einsum('gk,ik,jk->gij', w[None, :] ** (2 * wt[:, None]), np.abs(dc), np.abs(dc))
            assert (np.abs(G.cpu().numpy() - want) <= 1e-13 * scale).all()
        if top <= 3.0:
            want = orc.normvar(dt, dc, w, wt)
            got = nv.normvar(dt, dc, w, wt)
            np.testing.assert_allclose(got[0], want[0], rtol=1e-8, atol=1e-8 * np.abs(want[0]).max())


@gpu
def test_sym_pinv_matches_inv_rank():
    """nsr_sym_pinv (batched Jacobi) against the host inv_rank, rank-deficient cases included."""
    from normalisr_b200 import association, engine
    rng = np.random.default_rng(6)
    ctx = engine.context(0)
    for n in (1, 2, 5, 9, 12, 16):
        mats = []
        for k in range(40):
            a = rng.normal(size=(n, 3 * n + 2)) * rng.uniform(0.01, 100, size=(n, 1))
            if n > 2 and k % 3 == 0:
                a[1] = a[0] - 2 * a[n - 1]
            if n > 3 and k % 5 == 0:
                a[2] = 0
            mats.append(a @ a.T)
        G = torch.from_numpy(np.array(mats)).cuda()
        inv, rank = engine.sym_pinv(ctx, G)
        inv, rank = inv.cpu().numpy(), rank.cpu().numpy()
        for m, i, r in zip(mats, inv, rank):
            want, r0 = association.inv_rank(m)
            assert r == r0
            np.testing.assert_allclose(i, want, rtol=0, atol=1e-9 * np.abs(want).max())
