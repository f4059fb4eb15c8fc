"""The P-value routine of the CUDA epilogue (csrc/pvalue.cuh is host/device code) compiled for
the host and checked against the reference's scipy values and the mpmath tail points."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_p_close, load_golden


@pytest.fixture(scope="module")
def pv(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pv") / "libpvhost.so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "normalisr_b200", "csrc"),
                    os.path.join(ROOT, "tests", "pvalue_host_shim.cpp"), "-o", out], check=True, env=env)
    lib = ctypes.CDLL(out)
    lib.pv_each.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64]
    lib.pv_pairs.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64]
    return lib


def _each(lib, r2, a):
    r2 = np.ascontiguousarray(r2, dtype=np.float64).ravel()
    a = np.ascontiguousarray(np.broadcast_to(a, r2.shape), dtype=np.float64)
    out = np.empty_like(r2)
    lib.pv_each(r2.ctypes.data, a.ctypes.data, out.ctypes.data, r2.size)
    return out


def test_known_answer_grid(pv):
    g = load_golden("pvalue_kat")
    got = _each(pv, g["r2"], g["a"].ravel()).reshape(g["P"].shape)
    assert_p_close(got, g["P"], rtol=1e-9)


def test_random_grid_against_scipy(pv):
    from scipy.special import betainc
    rng = np.random.default_rng(1)
    a = 10 ** rng.uniform(np.log10(0.5), 6, 200000)
    r2 = 10 ** rng.uniform(-14, 0, 200000)
    assert_p_close(_each(pv, r2, a), betainc(a, 0.5, 1 - r2), rtol=1e-9)


def test_mpmath_tail_points(pv):
    pts = [(2000, 5, 0.05, 0.025494348631248), (100000, 5, 0.117, 1.1098053295708e-301),
           (10000, 5, 0.3, 5.6798583929825e-207), (1000000, 5, 0.037, 7.1939093799116e-300)]
    for n, c, r, want in pts:
        got = _each(pv, [r * r], (n - 1 - c) / 2)[0]
        assert abs(got - want) / want < 1e-9, (n, r, got, want)


def test_pair_version_is_bit_identical(pv):
    rng = np.random.default_rng(2)
    for a in (7.5, 15.0, 123.0, 4995.5, 49995.0):
        r2 = np.concatenate([10 ** rng.uniform(-14, 0, 5000), [0, 1, 1e-300, 0.3, 0.29999, 0.5]])
        r2 = np.ascontiguousarray(r2[:len(r2) // 2 * 2])
        one = _each(pv, r2, a)
        two = np.empty_like(r2)
        pv.pv_pairs(r2.ctypes.data, a, two.ctypes.data, r2.size)
        assert np.array_equal(one, two)
