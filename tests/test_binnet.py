"""binnet / bh (reference src/normalisr/binnet.py).  CPU: the oracle against the
golden vectors made by the unmodified reference, and the sort-free fixed-point rule the kernel
uses against the oracle.  GPU: the kernel through the public API, bit-identical booleans."""
import numpy as np
import pytest
import torch

import normalisr_oracle as orc
from conftest import load_golden


def _golden_nets(name):
    g = load_golden(name)
    P = g["P"]
    nets = [np.unpackbits(g["net"][k], axis=1)[:, :P.shape[1]].astype(bool) for k in range(len(g["qcut"]))]
    return P, g["qcut"], nets


def _oracle_net(P, q):
    try:
        return orc.binnet(P, q)
    except RuntimeError:
        return np.zeros(P.shape, dtype=bool)


@pytest.mark.parametrize("name", ["binnet_coex", "binnet_ties"])
def test_oracle_binnet_matches_reference(name):
    P, qcuts, nets = _golden_nets(name)
    for q, want in zip(qcuts, nets):
        assert np.array_equal(_oracle_net(P, q), want)


def test_oracle_bh_matches_reference():
    g = load_golden("bh_kat")
    for i in range(int(g["count"])):
        assert np.array_equal(orc.bh(g["p%d" % i]), g["q%d" % i])
    # weights: doubling every weight changes nothing; weight 2 == listing the value twice
    rng = np.random.default_rng(0)
    p = rng.random(50)
    np.testing.assert_allclose(orc.bh(p, np.full(50, 2.0)), orc.bh(p), rtol=1e-15)
    w = np.ones(50)
    w[7] = 2
    np.testing.assert_allclose(orc.bh(p, w), orc.bh(np.append(p, p[7]))[:50], rtol=1e-15)


def _fixed_point_row(p, qcut):
    """What csrc/binnet.cu does for one row (diagonal already removed)."""
    n0 = float(p.size)
    c = int((p <= qcut).sum())
    w = 1.0
    while c > 0:
        w = c / n0
        c_new = int((p / w <= qcut).sum())
        if c_new == c:
            break
        c = c_new
    return (p / w <= qcut) if c > 0 else np.zeros(p.size, dtype=bool)


def test_sort_free_rule_equals_bh_threshold():
    rng = np.random.default_rng(3)
    vals = np.array([0.0, 1e-300, 1e-9, 1e-4, 0.003, 0.01, 0.2, 0.5, 1.0])
    for trial in range(300):
        n = int(rng.integers(1, 400))
        kind = trial % 3
        if kind == 0:
            p = rng.random(n) ** rng.integers(1, 6)
        elif kind == 1:
            p = vals[rng.integers(0, vals.size, size=n)]
        else:                                        # adversarial staircase: many iterations
            p = np.sort(rng.random(n)) * np.arange(1, n + 1) / n * 0.3
        for q in (0.5, 0.05, 1e-3):
            assert np.array_equal(_fixed_point_row(p, q), orc.bh(p) <= q)


# ------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@pytest.fixture(params=["keys", "values"])
def row_kernel(request):
    """Both row kernels of csrc/binnet.cu: 4-byte keys in shared memory, two rows per SM (the default) and
    the 8-byte row staged with one bulk copy."""
    from normalisr_b200 import engine
    engine.set_option("binnet_keys", 1 if request.param == "keys" else 0)
    yield request.param
    engine.set_option("binnet_keys", 1)


@gpu
@pytest.mark.parametrize("name", ["binnet_coex", "binnet_ties"])
def test_binnet_golden_bit_identical(name, row_kernel):
    from normalisr_b200 import normalisr as norm
    P, qcuts, nets = _golden_nets(name)
    for q, want in zip(qcuts, nets):
        if want.sum() == 0:
            with pytest.raises(RuntimeError):
                norm.binnet(P, q)
            continue
        got = norm.binnet(P, q)
        assert isinstance(got, np.ndarray) and got.dtype == np.bool_
        assert np.array_equal(got, want)
        dev = norm.binnet(torch.from_numpy(P).cuda(), q)
        assert dev.is_cuda and dev.dtype == torch.bool and np.array_equal(dev.cpu().numpy(), want)


@gpu
def test_binnet_large_against_oracle_and_row_chunks(monkeypatch, row_kernel):
    """1,500 genes: coex P from the CUDA path, a hub row with > 6000... candidates is exercised at
    7,000 columns separately; host input staged in several row chunks."""
    from normalisr_b200 import binnet as bn, normalisr as norm, synth
    p = synth.host_problem(1005, 1500, 1200)
    P, _, _ = norm.coex(p["dt"], p["dc"])
    for q in (0.2, 1e-4):
        want = _oracle_net(P, q)
        assert want.sum() > 0
        assert np.array_equal(norm.binnet(P, q), want)
    monkeypatch.setattr(bn, "_ROW_CHUNK_BYTES", 8 * 1500 * 100)          # 100 rows per chunk
    assert np.array_equal(norm.binnet(P, 0.2), _oracle_net(P, 0.2))
    # dense rows (most entries significant) and a hub matrix
    rng = np.random.default_rng(8)
    W = rng.random((40, 7000)) ** 6
    W = np.concatenate([W, rng.random((7000 - 40, 7000))])              # square, 40 dense rows on top
    got = norm.binnet(W, 0.5)
    for i in (0, 17, 39, 40, 6999):
        row = np.delete(W[i], i)
        assert np.array_equal(np.delete(got[i], i), orc.bh(row) <= 0.5) and not got[i, i]


@gpu
def test_binnet_rows_wider_than_shared_memory(row_kernel):
    """Rows of more than 25,000 entries (100,000 for the key kernel) take the kernel variant that re-reads the row
    from L2; also rows of a row block whose diagonal sits at an offset, or outside the block."""
    from normalisr_b200 import binnet as bn, engine
    rng = np.random.default_rng(9)
    ctx = engine.context(0)
    for cols, diag0 in ((30011, 0), (30011, 29990), (30011, -5), (5000, 4990), (4999, 3), (24999, 100), (25000, 7), (49999, 3),
                        (50000, 11), (50001, 49000), (100003, 5)):
        Pm = rng.random((24, cols)) ** rng.integers(1, 8, size=(24, 1))
        out, stats = bn.binnet_rows(ctx, torch.from_numpy(Pm).cuda(), 0.1, diag0)
        got = out.cpu().numpy().astype(bool)
        edges = 0
        for i in range(24):
            d = i + diag0
            keep = np.ones(cols, dtype=bool)
            if 0 <= d < cols:
                keep[d] = False
                assert not got[i, d]
            want = orc.bh(Pm[i, keep]) <= 0.1
            assert np.array_equal(got[i, keep], want)
            edges += int(want.sum())
        assert int(stats[0]) == edges and int(stats[1]) == 0


@gpu
def test_binnet_slowly_converging_and_tied_rows(row_kernel):
    """Rows on which the plain fixed-point iteration needs many steps (a staircase just under the BH
    line), rows packed with ties (window overflow -> fallback path), negative zero, all-equal rows."""
    from normalisr_b200 import binnet as bn, engine
    rng = np.random.default_rng(10)
    ctx = engine.context(0)
    cols = 6000
    rows = []
    k = np.arange(1, cols + 1)
    rows.append(0.1 * k / cols * (1 - 1e-9 * rng.random(cols)))                  # every rank passes barely
    rows.append(0.1 * k / cols * (1 + 1e-3 * rng.random(cols)))                  # every rank fails barely
    rows.append(np.where(rng.random(cols) < 0.5, 0.0123456, rng.random(cols)))   # 3000 ties inside the window
    rows.append(np.full(cols, 0.04))
    rows.append(np.where(rng.random(cols) < 0.3, -0.0, rng.random(cols) ** 4))
    rows.append(np.exp(-rng.exponential(size=cols) * 40))
    rows.append(np.sort(rng.random(cols)) ** 2 * 0.2)
    Pm = np.array([rng.permutation(r) for r in rows])
    for q in (0.1, 0.05, 1e-3):
        out, stats = bn.binnet_rows(ctx, torch.from_numpy(Pm).cuda(), q, -cols - 10)   # no diagonal inside any row
        got = out.cpu().numpy().astype(bool)
        for i in range(len(rows)):
            assert np.array_equal(got[i], orc.bh(Pm[i]) <= q), (i, q)


@gpu
def test_binnet_threshold_ties_in_the_high_word(row_kernel):
    """Entries that share the high 32 bits of the BH threshold (the key kernel decides those from the value in
    global memory), rows at odd offsets (no 16-byte alignment), odd widths, a diagonal inside the cluster."""
    from normalisr_b200 import binnet as bn, engine
    rng = np.random.default_rng(12)
    ctx = engine.context(0)
    for cols, ld, off in ((4000, 4000, 0), (4001, 4001, 0), (4000, 4003, 1), (3999, 4002, 3)):
        rows = []
        for centre in (0.01, 0.05 * 1000 / cols, 3e-5):
            k = rng.integers(-2000, 2000, size=cols)
            cluster = centre * (1.0 + k * 2.0 ** -44)                      # equal high words, different low words
            rows.append(np.where(rng.random(cols) < 0.4, cluster, rng.random(cols) ** 3))
        rows.append(np.where(rng.random(cols) < 0.5, np.float64(0.05) * 1200 / cols, rng.random(cols)))
        Pm = np.array(rows)
        buf = torch.zeros(len(rows) * ld + 8, dtype=torch.float64, device="cuda")
        view = buf[off:off + len(rows) * ld].view(len(rows), ld)[:, :cols]
        view.copy_(torch.from_numpy(Pm))
        for q, diag0 in ((0.05, -cols - 5), (0.05, 7), (0.3, 1000)):
            out, stats = bn.binnet_rows(ctx, view, q, diag0)
            got = out.cpu().numpy().astype(bool)
            for i in range(len(rows)):
                d = i + diag0
                keep = np.ones(cols, dtype=bool)
                if 0 <= d < cols:
                    keep[d] = False
                    assert not got[i, d]
                assert np.array_equal(got[i, keep], orc.bh(Pm[i, keep]) <= q), (cols, ld, off, q, diag0, i)
            assert int(stats[1]) == 0


@gpu
def test_binnet_errors_and_bh():
    from normalisr_b200 import binnet as bn, normalisr as norm
    P = np.random.default_rng(1).random((50, 50))
    with pytest.raises(ValueError):
        norm.binnet(P[:, :40], 0.1)
    with pytest.raises(ValueError):
        norm.binnet(P[:1, :1], 0.1)
    for q in (0.0, 1.0, -1, 2):
        with pytest.raises(ValueError):
            norm.binnet(P, q)
    bad = P.copy()
    bad[3, 4] = np.nan
    with pytest.raises(AssertionError):
        norm.binnet(bad, 0.1)
    bad[3, 4] = 1.5
    with pytest.raises(AssertionError):
        norm.binnet(bad, 0.1)
    with pytest.raises(RuntimeError):
        norm.binnet(np.ones((20, 20)), 0.5)
    g = load_golden("bh_kat")
    for i in range(int(g["count"])):
        np.testing.assert_allclose(bn.bh(g["p%d" % i]), g["q%d" % i], rtol=1e-14, atol=0)
    w = np.random.default_rng(2).random(257) + 0.1
    np.testing.assert_allclose(bn.bh(g["p0"], w), orc.bh(g["p0"], w), rtol=1e-12)
    # nodiag / rediag round trip (binnet.py:4-74)
    m = np.arange(30.0).reshape(5, 6)
    assert np.array_equal(bn.rediag(bn.nodiag(m), fill=-1, shape=m.shape)[~np.eye(5, 6, dtype=bool)],
                          m[~np.eye(5, 6, dtype=bool)])
