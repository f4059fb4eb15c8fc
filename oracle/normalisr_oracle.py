"""CPU oracle for the Normalisr association-testing hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm
(lingfeiwang/normalisr v1.0.0, ``src/normalisr/association.py``, ``coex.py``,
``de.py``; and, for the rows of SURVEY 8(f) that were built, ``binnet.py``,
``norm.py:normvar`` / ``compute_var`` and ``lcpm.py:lcpm``).  It exists so that the CUDA path can be checked on a box where
``/root/reference`` is absent.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``normalisr_b200`` never does.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, run in the authoring container by ``tests/golden/make_golden*.py`` (coex / de
single 0 and 4, de single=1, binnet / bh, normvar, compute_var, lcpm) and committed as
``tests/golden/*.npz`` (``tests/test_oracle.py``, ``test_binnet.py``,
``test_normvar.py`` and ``test_lcpm.py`` check every one).

Third-party arithmetic the reference calls and that is not under /root/reference:
``numpy.matmul`` (BLAS dgemm), ``scipy.linalg.svd`` (LAPACK gesdd) and
``scipy.stats.beta.cdf`` -> ``scipy.special.betainc`` (unpinned in setup.py:29;
fixtures were made with numpy 2.3.5 / scipy 1.18.1).  matmul and svd are used
as-is from numpy; the incomplete beta function is used from scipy when it is
importable and otherwise (or with ``pvalue_backend='c'``) from the plain-C
restatement of its published algorithm in ``oracle/betainc_cf.c``.

All arrays follow the reference convention: float64, rows = variables
(genes / groupings / covariates), columns = cells.
"""
import ctypes
import itertools
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


# --------------------------------------------------------------------------------------
# P-value: regularised incomplete beta I_x(a, 1/2)       (association.py:249, 563)
# --------------------------------------------------------------------------------------
def _load_c():
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle C library missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.oracle_betainc_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                             ctypes.c_void_p, ctypes.c_int64]
        lib.oracle_betainc_array.restype = None
        _CLIB = lib
    return _CLIB


def beta_cdf(x, a, b=0.5, backend="auto"):
    """``scipy.stats.beta.cdf(x, a, b)`` as the reference calls it
    (association.py:249: ``beta.cdf(1 - R2, (n - 1 - dcr - dimreduce) / 2, 0.5)``).
    ``a`` may be a scalar or an array broadcastable against ``x``."""
    x = np.asarray(x, dtype=np.float64)
    if backend == "auto":
        try:
            import scipy.special  # noqa: F401
            backend = "scipy"
        except Exception:  # pragma: no cover
            backend = "c"
    if backend == "scipy":
        from scipy.special import betainc
        xc = np.clip(x, 0.0, 1.0)           # beta.cdf is 0 below and 1 above the support
        return betainc(a, b, xc)
    lib = _load_c()
    xa = np.ascontiguousarray(np.broadcast_to(x, np.broadcast(x, a).shape), dtype=np.float64)
    aa = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), xa.shape))
    out = np.empty_like(xa)
    lib.oracle_betainc_array(xa.ctypes.data, aa.ctypes.data, float(b), out.ctypes.data, xa.size)
    return out


# --------------------------------------------------------------------------------------
# inv_rank                                                (association.py:4-134)
# --------------------------------------------------------------------------------------
def pinv_rank(m, tol=1e-8, **unsupported):
    """SVD pseudo-inverse and rank of a square matrix; singular values below
    ``tol * s_max`` count as zero (association.py:66-80, the ``method='scipy'``
    branch with mpc == 0, which is the only one the hot path reaches by default)."""
    if unsupported.get("mpc", 0) or unsupported.get("method", "auto") not in ("auto", "scipy"):
        raise NotImplementedError("oracle covers the exact-SVD branch of inv_rank only")
    m = np.asarray(m, dtype=np.float64)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("Wrong shape for m.")
    if tol <= 0:
        raise ValueError("tol must be positive.")
    u, s, vt = np.linalg.svd(m)
    # number kept = n - #(s < tol*s0), exactly the searchsorted rule at :77
    keep = m.shape[0] - int(np.searchsorted(s[::-1], tol * s[0]))
    v = vt[:keep]
    return (v.T / s[:keep]) @ v, keep      # symmetric input, so this equals the reference's .T


# --------------------------------------------------------------------------------------
# association_test_1 : one tile, single=0                 (association.py:137-260)
# --------------------------------------------------------------------------------------
def tile_single0(dx, dy, dc, dci, dcr, dimreduce=0, lowmem=True, pvalue_backend="auto"):
    if dx.ndim != 2 or dy.ndim != 2 or dc.ndim != 2:
        raise ValueError("Incorrect dx/dy/dc size.")
    n = dx.shape[1]
    if dy.shape[1] != n or dc.shape[1] != n:
        raise ValueError("Unmatching dx/dy/dc dimensions.")
    if n <= dcr + dimreduce + 1:
        raise ValueError("Insufficient number of cells: must be greater than degrees of "
                         "freedom removed + covariate + 1.")
    rx, ry = dx, dy
    if dcr > 0:                                            # :224-229
        cx = (dci @ (dc @ dx.T)).T
        cy = (dci @ (dc @ dy.T)).T
        rx = dx - cx @ dc
        ry = dy - cy @ dc
    vx = np.mean(rx * rx, axis=1)                          # :230-233
    vx[vx == 0] = 1
    vy = np.mean(ry * ry, axis=1)
    vy[vy == 0] = 1
    gamma = (ry @ rx.T / (n * vx)).T                       # :234
    r2 = (gamma * gamma) * vx[:, None] / vy[None, :]       # :235
    if lowmem:
        alpha = None
    elif dcr > 0:                                          # :238-243  alpha = cy - gamma*cx
        alpha = cy[None, :, :] - gamma[:, :, None] * cx[:, None, :]
    else:
        alpha = np.zeros((dx.shape[0], dy.shape[0], dc.shape[0]))
    assert (r2 >= 0).all() and (r2 <= 1 + 1e-8).all()     # :248
    pv = beta_cdf(1 - r2, (n - 1 - dcr - dimreduce) / 2, 0.5, backend=pvalue_backend)  # :249
    return pv, gamma, alpha, vx, vy


# --------------------------------------------------------------------------------------
# association_test_4 : single=4 from Gram matrices        (association.py:421-576)
# --------------------------------------------------------------------------------------
def tile_single4(vx, prod, prody, prodyy, nx, nc, n, lenx, dimreduce=0, pvalue_backend="auto",
                 **ka):
    """dx != dy branch only (de never passes dy=None).  ``prod`` = A A^T with
    A = [dx; dc], ``prody`` = A dy^T, ``prodyy`` = rowsum(dy^2)."""
    ny = prody.shape[1]
    m = nx + nc
    pv = np.zeros((lenx, ny))
    gam = np.zeros((lenx, ny))
    varx = np.zeros(lenx)
    vary = np.zeros((lenx, ny))
    rank = np.zeros((lenx, ny), dtype=int)
    for i in range(lenx):
        x = vx + i
        oth = [k for k in range(m) if k != x]              # :523
        r = 0
        if oth:
            ginv, r = pinv_rank(prod[np.ix_(oth, oth)], **ka)   # :527-528
        rank[i] = r
        if r == 0:                                         # :532-536
            dxx = prod[x, x] / n
            dyy = prodyy / n
            dxy = prody[x] / n
        else:                                              # :537-544
            cx = prod[x, oth] @ ginv
            dxx = (prod[x, x] - cx @ prod[oth, x]) / n
            cy = prody[oth].T @ ginv
            dyy = (prodyy - np.sum(cy.T * prody[oth], axis=0)) / n
            dxy = (prody[x] - cy @ prod[oth, x]) / n
        if dxx == 0:                                       # :545-547
            dxx = 1
        varx[i] = dxx
        vary[i] = dyy
        gam[i] = dxy / dxx
        pv[i] = dxy * dxy / (dxx * dyy)
    assert (pv >= 0).all() and (pv <= 1 + 1e-8).all()      # :557
    dof = n - 1 - rank - dimreduce
    if (dof <= 0).any():
        raise RuntimeError("Insufficient number of cells: must be greater than degrees of "
                           "freedom removed + covariate + 1.")
    pv = beta_cdf(1 - pv, dof / 2, 0.5, backend=pvalue_backend)   # :563
    return pv, gam, varx, vary


def tile_single4_same(prod, nx, nc, n, dimreduce=0, lowmem=True, pvalue_backend="auto", **ka):
    """association_test_4 with dx == dy (``prody is None``, association.py:492-496, 517-519): every pair
    x < y is tested with ALL other rows of [dx; dc] as covariates - neither x nor y is one (:523-526).
    Returns (pv, gamma, vary, alpha|None) with entries for x < y only (zeros elsewhere)."""
    m = nx + nc
    pv = np.zeros((nx, nx))
    gam = np.zeros((nx, nx))
    vary = np.zeros((nx, nx))
    rank = np.zeros((nx, nx), dtype=int)
    alpha = None if lowmem else np.zeros((nx, nx, nc))
    for x in range(nx):
        for y in range(x + 1, nx):
            oth = [k for k in range(m) if k != x and k != y]
            r = 0
            if oth:
                ginv, r = pinv_rank(prod[np.ix_(oth, oth)], **ka)
            rank[x, y] = r
            if r == 0:
                dxx, dyy, dxy = prod[x, x] / n, prod[y, y] / n, prod[x, y] / n
            else:
                cx = prod[x, oth] @ ginv
                dxx = (prod[x, x] - cx @ prod[oth, x]) / n
                cy = prod[oth, y] @ ginv
                dyy = (prod[y, y] - cy @ prod[oth, y]) / n
                dxy = (prod[x, y] - cy @ prod[oth, x]) / n
            if dxx == 0:
                dxx = 1
            vary[x, y] = dyy
            gam[x, y] = dxy / dxx
            if not lowmem and r > 0 and nc > 0:                       # :551-553
                alpha[x, y] = cy[-nc:] - gam[x, y] * cx[-nc:]
            pv[x, y] = dxy * dxy / (dxx * dyy)
    assert (pv >= 0).all() and (pv <= 1 + 1e-8).all()
    dof = n - 1 - rank - dimreduce
    if (dof <= 0).any():
        raise RuntimeError("Insufficient number of cells: must be greater than degrees of "
                           "freedom removed + covariate + 1.")
    pv = beta_cdf(1 - pv, dof / 2, 0.5, backend=pvalue_backend)
    return pv, gam, vary, alpha


# --------------------------------------------------------------------------------------
# _auto_batchsize / association_tests                     (association.py:731-1093)
# --------------------------------------------------------------------------------------
def _batch(bs, itemsize, nc, ns, cap, sizemax=2 ** 30):
    if bs == 0:                                            # :743-746: two tiles <= 1 GiB
        bs = min(int((sizemax - itemsize * nc * ns) // (2 * itemsize * ns)), cap)
    return bs


def _blocks(n, bs):
    return [(s, min(s + bs, n)) for s in range(0, n, bs)]


def tile_single1(dx, dy, dc, sselectx, dimreduce=0, lowmem=True, pvalue_backend="auto"):
    """association_test_2 (association.py:263-390): like the single=0 tile, but every x is tested
    on its own subset of cells; covariates are projected out within that subset (own pseudo-inverse
    and rank per x)."""
    nx, n = dx.shape
    ny, nc = dy.shape[0], dc.shape[0]
    r2 = np.zeros((nx, ny))
    vx = np.zeros(nx)
    vy = np.zeros((nx, ny))
    gamma = np.zeros((nx, ny))
    alpha = None if lowmem else np.zeros((nx, ny, nc))
    rank = np.zeros((nx, ny), dtype=int)
    for xi in range(nx):
        sel = np.nonzero(sselectx[xi])[0]                       # :337-342
        ns = len(sel)
        if len(np.unique(dx[xi, sel])) < 2:
            continue
        x1, y1 = dx[xi, sel], dy[:, sel]
        r = 0
        if nc > 0:                                              # :343-350
            c1 = dc[:, sel]
            ci, r = pinv_rank(c1 @ c1.T)
        rank[xi] = r
        if r > 0:                                               # :352-357
            ccx = (ci @ (c1 @ x1.T)).T
            ccy = (ci @ (c1 @ y1.T)).T
            x1 = x1 - ccx @ c1
            y1 = y1 - ccy @ c1
        v = (x1 ** 2).mean()                                    # :358-362
        if v == 0:
            v = 1
        vx[xi] = v
        vy[xi] = (y1 ** 2).mean(axis=1)
        gamma[xi] = (x1 @ y1.T).ravel() / (ns * v)              # :364
        if not lowmem and r > 0:                                # :365-367
            alpha[xi] = ccy - gamma[xi][:, None] * ccx.ravel()
        r2[xi] = gamma[xi] ** 2 * v / vy[xi]                    # :368
    assert (r2 >= 0).all() and (r2 <= 1 + 1e-8).all()
    dof = (sselectx.sum(axis=1) - 1 - rank.T - dimreduce).T     # :372
    if (dof <= 0).any():
        raise RuntimeError('Insufficient number of cells: must be greater than degrees of freedom '
                           'removed + covariate + 1.')
    pv = beta_cdf(1 - r2, dof / 2, 0.5, backend=pvalue_backend)  # :377
    return pv, gamma, alpha, vx, vy


def association_tests(dx, dy, dc, bsx=0, bsy=0, nth=1, lowmem=True, return_dot=True, single=0,
                      bs4=500, pvalue_backend="auto", **ka):
    """Driver: tiling, per-tile kernels, assembly, gamma<->dot conversion and (for
    dy=None) symmetrisation.  ``nth`` > 1 maps tiles over a thread pool exactly like
    the reference's ``autopooler(..., dummy=True)`` (parallel.py:49-71)."""
    if single not in (0, 1, 4):
        raise ValueError("oracle covers single=0, 1 and 4")
    samexy = dy is None
    if samexy:
        dy = dx
        if single == 4:
            return _single4_same(dx, dc, lowmem, return_dot, pvalue_backend, **ka)
    nx, ns = dx.shape
    ny = dy.shape[0]
    nc = dc.shape[0]
    dimreduce = ka.pop("dimreduce", 0)
    capx, capy = (500, 500) if single in (0, 1) else (10, 500000)     # :854-875
    bsx = _batch(bsx, dx.dtype.itemsize, nc, ns, capx)
    bsy = bsx if samexy else _batch(bsy, dy.dtype.itemsize, nc, ns, capy)
    tiles = list(itertools.product(_blocks(nx, bsx), _blocks(ny, bsy)))
    if samexy:
        tiles = [t for t in tiles if t[0][0] <= t[1][0]]            # :892-893

    def pmap(fn, items):
        if nth == 1:
            return [fn(i) for i in items]
        from multiprocessing.dummy import Pool
        with Pool(nth if nth > 0 else os.cpu_count()) as p:
            return p.map(fn, items)

    P = np.ones((nx, ny))
    coef = np.zeros((nx, ny))
    alpha = None if lowmem else np.zeros((nx, ny, nc))
    varx = None if samexy else np.zeros(nx)
    if single == 0:
        if nc > 0 and (dc != 0).any():                              # :899-903
            dci, dcr = pinv_rank(dc @ dc.T)
        else:
            dci, dcr = None, 0
        vary = np.zeros(ny)

        def run(t):
            (x0, x1), (y0, y1) = t
            return t, tile_single0(dx[x0:x1], dy[y0:y1], dc, dci, dcr, dimreduce, lowmem,
                                   pvalue_backend)
        for ((x0, x1), (y0, y1)), (pv, g, al, vx, vy) in pmap(run, tiles):
            P[x0:x1, y0:y1] = pv
            coef[x0:x1, y0:y1] = g
            if not lowmem:
                alpha[x0:x1, y0:y1] = al
            if not samexy:
                varx[x0:x1] = vx
            vary[y0:y1] = vy
    elif single == 1:
        if samexy:
            raise NotImplementedError('dy=None with single=1')          # :911-912
        assert dx.max() == 1                                            # :914
        sselectx = dx == dx.sum(axis=0)                                 # :915-916
        for xi in range(nx):
            assert len(np.unique(dx[xi, sselectx[xi]])) > 1             # :917-918
        vary = np.zeros((nx, ny))

        def run1(t):
            (x0, x1), (y0, y1) = t
            return t, tile_single1(dx[x0:x1], dy[y0:y1], dc, sselectx[x0:x1], dimreduce, lowmem, pvalue_backend)
        for ((x0, x1), (y0, y1)), (pv, g, al, vx, vy) in pmap(run1, tiles):
            P[x0:x1, y0:y1] = pv
            coef[x0:x1, y0:y1] = g
            if not lowmem:
                alpha[x0:x1, y0:y1] = al
            varx[x0:x1] = vx
            vary[x0:x1, y0:y1] = vy
    else:
        A = np.concatenate([dx, dc], axis=0)                        # :935
        m = nx + nc
        prod = np.zeros((m, m))
        for (i0, i1), (j0, j1) in itertools.product(_blocks(m, bs4), _blocks(m, bs4)):
            if i0 <= j0:                                            # :936-950
                prod[i0:i1, j0:j1] = A[i0:i1] @ A[j0:j1].T
        prod = np.triu(prod).T + np.triu(prod, 1)
        prody = np.zeros((m, ny))
        for (i0, i1), (j0, j1) in itertools.product(_blocks(m, bs4), _blocks(ny, bs4)):
            prody[i0:i1, j0:j1] = A[i0:i1] @ dy[j0:j1].T            # :952-966
        prodyy = (dy ** 2).sum(axis=1)                              # :968
        vary = np.zeros((nx, ny))

        def run4(t):
            (x0, x1), (y0, y1) = t
            return t, tile_single4(x0, prod, prody[:, y0:y1], prodyy[y0:y1], nx, nc, ns, x1 - x0,
                                   dimreduce, pvalue_backend, **ka)
        for ((x0, x1), (y0, y1)), (pv, g, vx, vy) in pmap(run4, tiles):
            P[x0:x1, y0:y1] = pv
            coef[x0:x1, y0:y1] = g
            varx[x0:x1] = vx
            vary[x0:x1, y0:y1] = vy
    # gamma -> dot, symmetrise                                     :1036-1065
    if samexy:
        dot = coef * vary[:, None]            # row x times var of x (== vary for dy=dx)
        P = np.triu(P, 1)
        P = P + P.T
        dot = np.triu(dot, 1)
        dot = dot + dot.T
        if not return_dot:
            dot = dot / vary[:, None]
        if not lowmem:
            a = np.triu(alpha.transpose(2, 0, 1))
            alpha = (a + a.transpose(0, 2, 1)).transpose(1, 2, 0)
        coef = dot
    elif return_dot:
        coef = coef * varx[:, None]
    return P, coef, alpha, varx, vary


def _single4_same(dx, dc, lowmem, return_dot, pvalue_backend, **ka):
    """single=4 with dy=None: Gram matrix of [dx; dc] (:935-951), the pair loop, and the assembly of
    association.py:1036-1065 for this case (note :1040: the coefficient is multiplied by vary, not varx)."""
    nx, ns = dx.shape
    nc = dc.shape[0]
    dimreduce = ka.pop("dimreduce", 0)
    A = np.concatenate([dx, dc], axis=0)
    prod = A @ A.T
    pv, gam, vary, alpha = tile_single4_same(prod, nx, nc, ns, dimreduce, lowmem, pvalue_backend, **ka)
    dot = gam * vary                                   # :1040
    P = np.triu(pv, 1)
    P = P + P.T                                        # :1049-1050
    vary = np.triu(vary, 1)
    vary = vary + vary.T
    vary[np.arange(nx), np.arange(nx)] = 1             # :1051-1054
    dot = np.triu(dot, 1)
    dot = dot + dot.T                                  # :1055-1056
    if not return_dot:
        dot = dot / vary                               # :1061
    if not lowmem:
        a = np.triu(alpha.transpose(2, 0, 1))
        alpha = (a + a.transpose(0, 2, 1)).transpose(1, 2, 0)       # :1063-1065
    return P, dot, alpha, None, vary


# --------------------------------------------------------------------------------------
# coex / de                                                (coex.py:4-48, de.py:4-132)
# --------------------------------------------------------------------------------------
def coex(dt, dc, **ka):
    """(P, dot, var) for all gene pairs (coex.py:46-48)."""
    ans = association_tests(dt, None, dc, **ka)
    return ans[0], ans[1], ans[4]


def de(dg, dt, dc, bs=0, **ka):
    """(P, gamma, alpha|None, varg, vart); constant groupings are dropped and
    back-filled with P=1 / 0 (de.py:92-122)."""
    dg0 = np.asarray(dg)
    keep = np.array([len(np.unique(x)) > 1 for x in dg0], dtype=bool)
    P, gam, alpha, vg, vt = association_tests(dg0[keep], dt, dc, bsx=bs, bsy=bs,
                                              return_dot=False, **ka)
    ng, nt, nc = dg0.shape[0], dt.shape[0], dc.shape[0]
    Pf = np.ones((ng, nt), dtype=dt.dtype)
    Pf[keep] = P
    gf = np.zeros((ng, nt), dtype=dt.dtype)
    gf[keep] = gam
    af = None
    if alpha is not None:
        af = np.zeros((ng, nt, nc), dtype=dt.dtype)
        af[keep] = alpha
    vgf = np.zeros(ng, dtype=dt.dtype)
    vgf[keep] = vg
    vtf = np.zeros((ng, nt), dtype=dt.dtype)
    vtf[keep] = vt                                  # (nt,) broadcasts for single=0 (:120-121)
    return Pf, gf, af, vgf, vtf


# --------------------------------------------------------------------------------------
# next row of SURVEY 8(f): P-value network -> per-row BH Q-values -> binary network
# (reference src/normalisr/binnet.py:77-173)
# --------------------------------------------------------------------------------------
def bh(pv, weight=None):
    """Benjamini-Hochberg Q-values, binnet.py:77-131: unique P-values, cumulative (weighted)
    counts normalised to 1, q = p / w clipped to [0, 1], running minimum from the largest P."""
    pv = np.asarray(pv)
    assert pv.ndim == 1 and pv.size > 0
    assert np.isfinite(pv).all() and pv.min() >= 0 and pv.max() <= 1
    weight = np.ones(pv.size) if weight is None else np.asarray(weight)
    assert weight.shape == pv.shape
    vals, inv = np.unique(pv, return_inverse=True)                 # :113
    w = np.zeros(vals.size, dtype=pv.dtype)
    np.add.at(w, inv, weight)                                      # :117-119 (same order of additions)
    w = np.cumsum(w)                                               # :122
    w /= w[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        q = vals / w                                               # :124
    q[~np.isfinite(q)] = 1
    q = np.minimum(q, 1)
    q = np.maximum(q, 0)
    q = np.minimum.accumulate(q[::-1])[::-1]                       # :128-129
    return q[inv].astype(pv.dtype, copy=False)


def binnet(net, qcut):
    """binnet.py:134-170: BH per row over the off-diagonal entries, diagonal Q = 1, threshold."""
    net = np.asarray(net)
    assert net.ndim == 2 and np.isfinite(net).all() and net.min() >= 0 and net.max() <= 1
    nt = net.shape[0]
    if net.shape[1] != nt or nt <= 1:
        raise ValueError('Wrong shape of net or namet.')
    if qcut <= 0 or qcut >= 1:
        raise ValueError('Q-value cutoff must be between 0 and 1.')
    q = np.ones_like(net)
    off = ~np.eye(nt, dtype=bool)
    for i in range(nt):
        q[i, off[i]] = bh(net[i, off[i]])                          # nodiag / rediag, :156-157
    ans = q <= qcut
    if ans.sum() == 0:
        raise RuntimeError("Empty binary network.")
    return ans


# --------------------------------------------------------------------------------------
# variance normalisation, the step directly upstream of coex / de
# (reference src/normalisr/norm.py:131-289)
# --------------------------------------------------------------------------------------
def normvar1(dt, dc, w2=None):
    """norm.py:131-166: remove covariates from every row of dt; with w2 (n_gene, n_cell) each
    gene uses its own covariates dc * w2[gene]."""
    if w2 is not None:
        return np.concatenate([normvar1(dt[[x]], dc * w2[x]) for x in range(dt.shape[0])], axis=0)
    g = dc @ dc.T                                                   # :159
    gi, r = pinv_rank(g)
    if r <= 0:
        raise RuntimeError('Zero-rank covariates found.')
    out = dt - (dc.T @ (gi @ (dc @ dt.T))).T                        # :163
    assert np.isfinite(out).all()
    return out


def normvar(dt, dc, w, wt, dextra=None, cat=1, keepvar=True, normmean=False, **ignored):
    """norm.py:169-289: dt * w**wt[gene], covariates dc * w**wt[gene] removed per gene, variance
    restored (keepvar), continuous covariates scaled by w."""
    if any(x.ndim != 2 for x in (dt, dc)):
        raise ValueError('dt and dc should have 2 dimensions.')
    if any(x.ndim != 1 for x in (w, wt)):
        raise ValueError('w and wt should have 1 dimension.')
    nt, ns = dt.shape
    if dc.shape[0] == 0:
        raise ValueError('No covariates.')
    if dc.shape[1] != ns or w.shape[0] != ns or wt.shape[0] != nt:
        raise ValueError('Unmatched gene or cell counts.')
    if dextra is not None and (dextra.ndim != 2 or dextra.shape[0] == 0 or dextra.shape[1] != ns):
        raise ValueError('Unmatched shape or size for dextra.')
    if w.min() <= 0:
        raise ValueError('w must be positive.')
    if wt.min() < 0:
        raise ValueError('wt must be non-negative.')
    w2 = (np.repeat([w], nt, axis=0).T ** wt).T                     # :238
    w2[wt == 0] = 1
    dt = dt * w2
    if keepvar:                                                     # :241-243
        dv = dt.mean(axis=1)
        dv = np.sqrt(((dt.T - dv) ** 2).mean(axis=0))
    dtn = normvar1(dt, dc, w2)                                      # :245-249
    if keepvar:                                                     # :251-254
        dv2 = np.sqrt((dtn ** 2).mean(axis=1))
        dtn = (dtn.T * ((dv / dv2) ** wt)).T
    if cat == 2:                                                    # :257-269
        dcn = dc * w
    elif cat == 1:
        dcn = dc.copy()
        t0 = ((dc != 0) & (dc != 1)).any(axis=1) | ((dc == 1).all(axis=1))
        dcn[t0] = dc[t0] * w
    elif cat == 0:
        dcn = dc.copy()
        t0 = ((dc != 0) & (dc != 1)).any(axis=1)
        dcn[t0] = dc[t0] * w
    else:
        raise ValueError('Invalid cat value.')
    if normmean:                                                    # :271-273
        dtn = normvar1(dtn, dcn)
    ans = [dtn, dcn]
    if dextra is not None:
        ans.append(dextra * w)
    return ans


# --------------------------------------------------------------------------------------
# Bayesian logCPM (reference src/normalisr/lcpm.py:21-208)
# --------------------------------------------------------------------------------------
def lcpm(reads, normalize=True, ntot=None, lowmem=True, nocov=False, varscale=0, noise=None, **ignored):
    """(lcpm, mean|None, var|None, cov|None): digamma(1 + reads) - digamma(total + 2), per-cell
    log-sum-exp normalisation to log counts per million, cellular covariates.  varscale != 0
    (posterior resampling, :104-109, 134-150, 178-190): value = mean + sqrt(varscale * (trigamma(1 +
    reads) - trigamma(total + 2))) * z with z = ``noise`` (the standard normal deviates; the reference
    draws numpy.random.randn(n_gene, n_cell) once)."""
    from scipy.special import digamma, polygamma
    d = np.asarray(reads)
    if d.ndim != 2:
        raise ValueError('reads must have 2 dimensions.')
    if (d < 0).any():
        raise ValueError('Negative value in d detected.')
    if not np.issubdtype(d.dtype, np.integer):
        d = d.astype(int)
    nt, nc = d.shape
    t0 = d.sum() + 2 if ntot is None else ntot + 2                   # :90
    assert t0 > 2                                                    # :91
    table = digamma(1.0 + np.arange(int(d.max()) + 1)) - float(digamma(t0))   # :96-109
    dmean = table[d]                                                 # :150, :176
    dvar = np.zeros(dmean.shape)
    if varscale != 0:
        tvar = polygamma(1, 1.0 + np.arange(int(d.max()) + 1)) - float(polygamma(1, t0))   # :104-109
        dvar = tvar[d] * varscale                                    # :143, :177-180
        dtn = np.asarray(noise) * np.sqrt(dvar) + dmean              # :144-147, :181-182
    else:
        dtn = dmean.copy()
    if normalize:                                                    # :155-157, :186-190
        shift = np.log(np.exp(dtn).sum(axis=0)) - np.log(1e6)
        dtn = dtn - shift
        dmean = dmean - shift
    dcov = None
    if not nocov:                                                    # :193-199
        t1 = d.sum(axis=0)
        if (t1 == 0).any():
            raise ValueError('Found cell with no read at all. Please remove.')
        t1 = np.log(t1)
        dcov = np.array([t1, nt - (d != 0).sum(axis=0), t1 ** 2])
    if lowmem:
        return dtn, None, None, dcov
    return dtn, dmean, dvar, dcov


def compute_var(dt, dc, stepmax=1, eps=1e-6):
    """norm.py:56-128: per-cell variance normalisation multiplier from a log-linear fit of each
    cell's residual variance on the covariates (EM-like iterations for stepmax > 1).  The
    reference fits with sklearn LinearRegression, i.e. least squares (minimum norm), without and
    with intercept."""
    if eps <= 0 or stepmax <= 0:
        raise ValueError('eps and stepmax must be positive.')
    if dt.ndim != 2 or dc.ndim != 2:
        raise ValueError('dt and dc must both have 2 dimensions.')
    if dt.shape[1] != dc.shape[1]:
        raise ValueError('dt and dc must have the same cell count.')
    ns = dc.shape[1]
    scale = np.ones(ns)
    best, bestv, n = None, 1e300, 0
    while n < stepmax and bestv > eps:
        td1, tdx = dt / scale, dc / scale                                     # :96-97
        beta = np.linalg.lstsq(tdx.T, td1.T, rcond=None)[0]                   # :98 (no intercept)
        td1 = td1 - (tdx.T @ beta).T                                          # :99
        m = td1.mean(axis=1)                                                  # :101
        sd = np.sqrt(((td1.T - m) ** 2).mean(axis=0))                         # :102
        y = np.log(np.sqrt((((td1.T - m) / sd) ** 2).mean(axis=1)))           # :103
        xc = dc.T - dc.T.mean(axis=0)                                         # :104-105 (with intercept)
        b2 = np.linalg.lstsq(xc, y - y.mean(), rcond=None)[0]
        new = np.exp(y.mean() + xc @ b2) * scale                              # :107
        new /= new.min()
        t1 = np.abs((new - scale) / scale).max()
        scale = new
        n += 1
        if t1 < bestv:
            bestv, best = t1, scale
    w = 1 / best
    return w / w.min()
