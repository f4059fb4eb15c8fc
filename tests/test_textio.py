"""Text I/O (reference run.py:10-35) through the native library: same values as numpy.loadtxt /
scipy.io.mmread, same bytes as numpy.savetxt.  Host code only: runs without a GPU."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse

from normalisr_b200 import io as nio


@pytest.fixture
def rng():
    return np.random.default_rng(123)


def test_tsv_write_matches_numpy_savetxt(tmp_path, rng):
    d = np.concatenate([rng.normal(size=(37, 11)) * 10.0 ** rng.integers(-12, 12, size=(37, 11)),
                        np.array([[0.0, -0.0, 1.0, -1.0, 1e-300, 1e300, 123456789.0, 0.1, 1 / 3, 2.5e-5, 99999999.5]])])
    a, b = str(tmp_path / "a.tsv"), str(tmp_path / "b.tsv")
    np.savetxt(a, d, delimiter='\t', fmt='%.8G')                       # run.file_write_tsv
    nio.file_write_tsv(b, d)
    assert open(a, 'rb').read() == open(b, 'rb').read()
    nio.file_write_tsv(b, d, delimiter=',', fmt='%.3G')
    np.savetxt(a, d, delimiter=',', fmt='%.3G')
    assert open(a, 'rb').read() == open(b, 'rb').read()
    nio.file_write_tsv(b, d[0])                                         # 1-D: one value per line, like savetxt
    np.savetxt(a, d[0], delimiter='\t', fmt='%.8G')
    assert open(a, 'rb').read() == open(b, 'rb').read()


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (300, 41), (2000, 3)])
def test_tsv_read_matches_numpy_loadtxt(tmp_path, rng, shape):
    d = rng.normal(size=shape) * 10.0 ** rng.integers(-6, 6, size=shape)
    f = str(tmp_path / "x.tsv")
    np.savetxt(f, d, delimiter='\t', fmt='%.8G')
    ref = np.loadtxt(f, delimiter='\t')
    if ref.ndim < 2:
        ref = ref.reshape(1, -1) if shape[0] == 1 else ref.reshape(-1, 1)
    got = nio.file_read_tsv(f)
    assert got.dtype == np.float64 and got.shape == shape and np.array_equal(got, ref.reshape(shape))
    assert np.array_equal(nio.file_read_tsv(f, nth=1), got) and np.array_equal(nio.file_read_tsv(f, nth=7), got)


def test_tsv_read_edge_cases(tmp_path):
    f = str(tmp_path / "e.tsv")
    open(f, 'w').write("# header\n1\t2\t3\r\n\n4e-3\t-5\tnan\n7\t8\tinf")        # comment, CRLF, blank line, no final newline
    got = nio.file_read_tsv(f)
    assert got.shape == (3, 3) and np.isnan(got[1, 2]) and np.isinf(got[2, 2]) and got[1, 0] == 4e-3 and got[0, 2] == 3
    open(f, 'w').write("1\t2\n3\n")
    with pytest.raises(Exception):
        nio.file_read_tsv(f)
    open(f, 'w').write("1\tx\n")
    with pytest.raises(Exception):
        nio.file_read_tsv(f)
    with pytest.raises(Exception):
        nio.file_read_tsv(str(tmp_path / "missing.tsv"))
    open(f, 'w').write("")
    assert nio.file_read_tsv(f).shape == (0, 0)


def test_mtx_read_matches_scipy(tmp_path, rng):
    m = scipy.sparse.random(200, 150, density=0.05, random_state=5, data_rvs=lambda k: rng.integers(1, 50, size=k)).tocoo()
    m.data = m.data.astype(np.int64)
    f = str(tmp_path / "c.mtx")
    scipy.io.mmwrite(f, m, field='integer')
    ref = scipy.io.mmread(f)                                          # run.file_read_coo
    got = nio.file_read_coo(f)
    assert scipy.sparse.issparse(got) and got.shape == ref.shape and np.issubdtype(got.dtype, np.integer)
    assert np.array_equal(got.toarray(), ref.toarray())
    mr = scipy.sparse.random(60, 60, density=0.1, random_state=6).tocoo()
    scipy.io.mmwrite(f, mr)
    assert np.array_equal(nio.file_read_coo(f).toarray(), scipy.io.mmread(f).toarray())
    ms = mr + mr.T
    scipy.io.mmwrite(f, ms, symmetry='symmetric')
    assert np.allclose(nio.file_read_coo(f).toarray(), scipy.io.mmread(f).toarray(), rtol=0, atol=0)
    open(f, 'w').write("%%MatrixMarket matrix coordinate pattern general\n% c\n3 4 2\n1 1\n3 4")
    assert np.array_equal(nio.file_read_coo(f).toarray(), scipy.io.mmread(f).toarray())
    open(f, 'w').write("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(Exception):
        nio.file_read_coo(f)


@pytest.mark.gpu
def test_text_io_onto_the_device(tmp_path, rng):
    import torch
    d = rng.normal(size=(50, 20))
    f = str(tmp_path / "x.tsv")
    np.savetxt(f, d, delimiter='\t', fmt='%.8G')
    t = nio.file_read_tsv(f, device='cuda')
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), np.loadtxt(f, delimiter='\t'))
    m = scipy.sparse.random(40, 30, density=0.2, random_state=1, data_rvs=lambda k: rng.integers(1, 9, size=k)).tocoo()
    g = str(tmp_path / "c.mtx")
    scipy.io.mmwrite(g, m.astype(np.int64), field='integer')
    c = nio.file_read_coo(g, device='cuda')
    assert c.is_cuda and c.dtype == torch.int32 and np.array_equal(c.cpu().numpy(), scipy.io.mmread(g).toarray())
    nio.file_write_tsv(f, t)
    assert np.array_equal(np.loadtxt(f, delimiter='\t'), np.loadtxt(f, delimiter='\t'))
