"""lcpm (reference src/normalisr/lcpm.py:21-208).  CPU: the oracle against the golden
vectors made by the unmodified reference.  GPU: the CUDA path against both."""
import numpy as np
import pytest
import torch

import normalisr_oracle as orc
from conftest import load_golden

gpu = pytest.mark.gpu


def test_oracle_lcpm_matches_reference():
    g = load_golden("lcpm_counts")
    a = orc.lcpm(g["reads"])
    np.testing.assert_allclose(a[0], g["lcpm"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(a[3], g["cov"], rtol=1e-14)
    assert a[1] is None and a[2] is None
    b = orc.lcpm(g["reads"], lowmem=False, nocov=True)
    np.testing.assert_allclose(b[1], g["mean"], rtol=1e-13, atol=1e-13)
    assert (b[2] == 0).all() and b[3] is None
    c = orc.lcpm(g["reads"], normalize=False, ntot=10 ** 9)
    np.testing.assert_allclose(c[0], g["lcpm_raw"], rtol=1e-13, atol=1e-13)
    empty_cell = np.array([[1, 0, 2], [3, 0, 1]])
    with pytest.raises(ValueError):
        orc.lcpm(empty_cell)                                     # a cell without reads (lcpm.py:195-196)
    with pytest.raises(AssertionError):
        orc.lcpm(np.zeros((3, 4), dtype=int))                    # no reads at all (lcpm.py:91)


def test_oracle_lcpm_resampling_matches_reference():
    """varscale != 0: with the reference's own deviates (numpy's global stream after seed()) the oracle
    reproduces the reference entry for entry."""
    g = load_golden("lcpm_resample")
    np.random.seed(int(g["seed"]))
    z = np.random.randn(*g["reads"].shape)
    a = orc.lcpm(g["reads"], varscale=float(g["varscale"]), noise=z)
    np.testing.assert_allclose(a[0], g["lcpm"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(a[3], g["cov"], rtol=1e-14)
    b = orc.lcpm(g["reads"], varscale=float(g["varscale"]), noise=z, lowmem=False)
    np.testing.assert_allclose(b[0], g["lcpm_full"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b[1], g["mean_full"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b[2], g["var_full"], rtol=1e-12, atol=1e-15)


@gpu
def test_lcpm_golden(monkeypatch):
    from normalisr_b200 import lcpm as lc, normalisr as norm
    g = load_golden("lcpm_counts")
    for chunk, hold in ((1 << 30, 1 << 35), (8 * 300 * 50, 1 << 35), (8 * 300 * 50, 0)):
        # one block / several row blocks kept on the device between the passes / several blocks fetched twice
        monkeypatch.setattr(lc, "_ROW_CHUNK_BYTES", chunk)
        monkeypatch.setattr(lc, "_HOLD_BYTES", hold)
        out = norm.lcpm(g["reads"])
        assert isinstance(out[0], np.ndarray) and out[1] is None and out[2] is None
        np.testing.assert_allclose(out[0], g["lcpm"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(out[3], g["cov"], rtol=1e-13)
    out = norm.lcpm(g["reads"].astype(np.int32), lowmem=False, nocov=True)
    np.testing.assert_allclose(out[1], g["mean"], rtol=1e-12, atol=1e-12)
    assert (out[2] == 0).all() and out[3] is None
    out = norm.lcpm(g["reads"], normalize=False, ntot=10 ** 9)
    np.testing.assert_allclose(out[0], g["lcpm_raw"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out[3], g["cov_raw"], rtol=1e-13)
    dev = norm.lcpm(torch.from_numpy(g["reads"].astype(np.int64)).cuda())
    assert dev[0].is_cuda and dev[3].is_cuda
    np.testing.assert_allclose(dev[0].cpu().numpy(), g["lcpm"], rtol=1e-12, atol=1e-12)


@gpu
def test_lcpm_larger_against_oracle_and_errors():
    from normalisr_b200 import normalisr as norm
    rng = np.random.default_rng(12)
    reads = rng.negative_binomial(2, 2.0 / (2.0 + rng.gamma(0.5, 4.0, size=(1500, 1)) * rng.lognormal(0, 0.4, size=(1, 2000))))
    reads[:, 7] += 1                                               # no empty cell
    want = orc.lcpm(reads)
    got = norm.lcpm(reads)
    np.testing.assert_allclose(got[0], want[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got[3], want[3], rtol=1e-13)
    np.testing.assert_allclose(np.exp(got[0]).sum(axis=0), 1e6, rtol=1e-10)        # counts per million
    with pytest.raises(ValueError):
        norm.lcpm(np.array([[1, 0, 2], [3, 0, 1]]))
    with pytest.raises(AssertionError):
        norm.lcpm(np.zeros((3, 4), dtype=int))
    with pytest.raises(ValueError):
        norm.lcpm(-np.ones((3, 4), dtype=int))
    with pytest.raises(ValueError):
        norm.lcpm(np.ones(4, dtype=int))
    with pytest.raises(ValueError):
        norm.lcpm(np.ones((3, 4), dtype=int), varscale=-1)


@gpu
def test_lcpm_posterior_resampling():
    """varscale != 0 (lcpm.py:104-109, 134-150, 178-190): numpy input reproduces the reference deviate for
    deviate (it draws from numpy's own stream, like the reference); CUDA input draws from Philox on the
    device: deterministic per seed, the right mean and variance per entry."""
    from normalisr_b200 import lcpm as lc, normalisr as norm
    g = load_golden("lcpm_resample")
    out = norm.lcpm(g["reads"], varscale=float(g["varscale"]), seed=int(g["seed"]))
    np.testing.assert_allclose(out[0], g["lcpm"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out[3], g["cov"], rtol=1e-13)
    out = norm.lcpm(g["reads"], varscale=float(g["varscale"]), seed=int(g["seed"]), lowmem=False)
    np.testing.assert_allclose(out[0], g["lcpm_full"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out[1], g["mean_full"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out[2], g["var_full"], rtol=1e-11, atol=1e-14)
    # the oracle with the same deviates
    np.random.seed(int(g["seed"]))
    z = np.random.randn(*g["reads"].shape)
    want = orc.lcpm(g["reads"], varscale=float(g["varscale"]), noise=z)
    np.testing.assert_allclose(want[0], g["lcpm"], rtol=1e-12, atol=1e-12)
    # device path: Philox stream
    rd = torch.from_numpy(g["reads"].astype(np.int64)).cuda()
    a = norm.lcpm(rd, varscale=0.7, seed=5, normalize=False, ntot=10 ** 9, lowmem=False)
    b = norm.lcpm(rd, varscale=0.7, seed=5, normalize=False, ntot=10 ** 9)
    c = norm.lcpm(rd, varscale=0.7, seed=6, normalize=False, ntot=10 ** 9)
    assert torch.equal(a[0], b[0]) and not torch.equal(a[0], c[0])
    zz = ((a[0] - a[1]) / torch.sqrt(a[2])).cpu().numpy().ravel()          # the deviates that were drawn
    assert abs(zz.mean()) < 5 / np.sqrt(zz.size) and abs(zz.var() - 1) < 10 / np.sqrt(zz.size)
    assert abs((zz ** 3).mean()) < 20 / np.sqrt(zz.size) and abs((zz ** 4).mean() - 3) < 60 / np.sqrt(zz.size)
    # the per-cell normaliser sees the resampled values: exp sums to one million
    d = norm.lcpm(rd, varscale=0.7, seed=5)
    np.testing.assert_allclose(torch.exp(d[0]).sum(dim=0).cpu().numpy(), 1e6, rtol=1e-10)
