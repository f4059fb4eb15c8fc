"""single=4: test every grouping x against every gene y with all OTHER groupings and the
covariates as nuisance regressors (reference ``association_test_4``, association.py:421-576,
driven from ``association_tests`` :926-974).

The reference forms the Gram matrices of A = [dx; dc] (``prod1`` tiles) and then, for every
x, pseudo-inverts the (m-1)x(m-1) Gram matrix of everything but x.  Algebraically that is the
multiple regression of y on all rows of A, so here:

  1. dx and dy are residualised against dc by the projection kernels (float64);
  2. Gxy = Rx Ry^T comes from the tensor-core contraction in NSR_MODE_RAW (exact integer sums;
     3 digit products when the groupings are small integers and travel as one exact plane);
     Gxx = Rx Rx^T is formed in float64 from the raw groupings like the reference's own prod1 tiles
     (exact integer sums on the tensor cores for small-integer groupings, ``nsr_gram_f64`` otherwise)
     minus the covariate part Cx Cx^T;
  3. ``nsr_de4_solve``: one blocked Cholesky factorisation of Gxx on the device gives every
     leave-one-out quantity in closed form (Schur complements), fused with the P-value.

When a Cholesky pivot falls below tol * max diag (numerically rank-deficient groupings) the closed
form does not apply and step 3 falls back to the reference's per-x pseudo-inverse, on the same Gram
matrices (``_loo_pinv``, host; the reference itself fails its range assert on exactly collinear
groupings, association.py:557).

Deliberate deviations from the reference (documented, both in unusual inputs only):
  * the rank rule is applied to the covariate-residualised groupings' Gram matrix, not to the joint
    Gram matrix of [other groupings; covariates] (association.py:527-528): covariates that are not
    scaled to O(1) (normcov scales them) can make the reference drop grouping directions that are kept
    here;
  * a gene that the covariates explain exactly has var 0 -> 1 in the projection (association.py:231) and
    gets P = 1 here, where the reference's single=4 branch divides by dyy = 0 and fails its assert.
"""
import logging

import numpy as np
import torch

from . import engine
from ._lib import MODE_RAW, MAX_RANK


def _exact_inverse_only(ka, size):
    """The keyword arguments single=4 hands to inv_rank (association.py:528: ``method``, ``mpc``, ``qr``).  ``mpc``
    exists to make the reference's per-grouping SVD affordable by truncating it; here every leave-one-out inverse
    comes from one factorisation, so only the settings under which the reference computes the exact pseudo-inverse
    are taken: method 'auto' / 'scipy' with mpc = 0 or mpc >= the size of the matrices inverted (then :64 picks
    the exact SVD and :78-79 truncates nothing).  ``qr`` only concerns the randomised SVD."""
    mpc = ka.pop('mpc', 0)
    method = ka.pop('method', 'auto')
    ka.pop('qr', None)
    if method not in ('auto', 'scipy') or (mpc != 0 and mpc < size):
        raise NotImplementedError('normalisr_b200 computes the exact leave-one-out inverses of single=4; a truncated '
                                  'or randomised SVD (method={!r}, mpc={} < {}) is not reproduced.'.format(method, mpc, size))


def association_tests_single4(dx, dy, dc, lowmem=True, return_dot=True, dimreduce=0,
                              precision='default', device=None, engine_id=None, exact_groupings=True, **ka):
    from .association import covariate_basis_device, _residualize_any, _residualize_groupings, _is_dev, _outs
    eng = ka.pop('engine', engine.ENGINE_UMMA) if engine_id is None else engine_id
    tol = ka.pop('tol', 1e-8)
    _exact_inverse_only(ka, dx.shape[0] - 1 + dc.shape[0])
    if ka:
        raise TypeError("association_test_4() got an unexpected keyword argument '{}'".format(
            next(iter(ka))))
    if tol <= 0:
        raise ValueError('tol must be positive.')
    to_host = not _is_dev(dx)
    ctx = engine.context(device if device is not None else (dx.device if _is_dev(dx) else None))
    n_slices, n_products = engine.PRESETS[precision]
    nx, n = dx.shape
    ny, nc = dy.shape[0], dc.shape[0]
    if nx == 0 or ny == 0 or n == 0:
        raise ValueError('Dimensions in na==0 detected.')
    if nc == 0:
        logging.warning('No covariate dc input.')
    Qt_dev, rank_c, W = covariate_basis_device(ctx, dc, tol=tol)
    if rank_c > MAX_RANK:
        raise NotImplementedError('covariate rank {} > {}'.format(rank_c, MAX_RANK))
    if n - 1 - (nx - 1 + rank_c) - dimreduce <= 0:
        raise RuntimeError('Insufficient number of cells: must be greater than degrees of '
                           'freedom removed + covariate + 1.')
    with torch.cuda.device(ctx.device):
        dev = ctx.device
        Rx, xd = _residualize_groupings(ctx, dx, Qt_dev, n_slices, True, exact=exact_groupings)
        Ry = _residualize_any(ctx, dy, Qt_dev, n_slices, not lowmem)
        Gxy = torch.empty((nx, ny), dtype=torch.float64, device=dev)
        engine.contract(ctx, MODE_RAW, Rx, Ry, engine.rect_tiles(nx, ny), 1.0, None, Gxy, n_products, eng)
        if Rx.n_slices == 1:
            # small-integer groupings: raw Gram matrix as exact integer sums (1 digit product), then - Cx Cx^T
            Gxx = torch.empty((nx, nx), dtype=torch.float64, device=dev)
            engine.contract(ctx, MODE_RAW, Rx, Rx, engine.rect_tiles(nx, nx), 1.0, None, Gxx, 1, eng)
            engine.gram_correct(ctx, Gxx, Rx.coef)
        else:
            Gxx = engine.gram_f64(ctx, xd, Rx.coef)
        yy = _row_sumsq(Ry)
        Gkeep = Gxx.clone()
        P, out2, dyy, dxx, w, status = engine.de4_solve(ctx, Gxx, Gxy, yy, n, rank_c, dimreduce, tol, return_dot)
        st = int(status.item())
        if st & 1:
            # numerically rank-deficient groupings: the reference's per-x pseudo-inverse (host)
            dxx, dxy, dyy, rank, w = _loo_pinv(Gkeep, Gxy, yy, n, rank_c, tol)
            dxx = torch.where(dxx == 0, torch.ones_like(dxx), dxx)               # :545-547
            gamma = dxy / dxx[:, None]
            r2 = dxy * dxy / (dxx[:, None] * dyy)
            if not bool(((r2 >= 0) & (r2 <= 1 + 1e-8)).all()):                   # :557
                raise AssertionError('R^2 outside [0, 1]: collinear groupings?')
            dof = n - 1 - rank - dimreduce
            if bool((dof <= 0).any()):
                raise RuntimeError('Insufficient number of cells: must be greater than degrees of '
                                   'freedom removed + covariate + 1.')
            P = engine.pvalue(ctx, r2, dof / 2)
            out2 = gamma * dxx[:, None] if return_dot else gamma
        elif st & 2:
            raise AssertionError('R^2 outside [0, 1] or non-finite result: collinear groupings?')       # :557, :565-568
        alpha = None
        if not lowmem:
            if rank_c:
                Wd = torch.from_numpy(W).to(dev)
                al = Ry.coef @ Wd - w.T @ (Rx.coef @ Wd)                      # (ny, nc)
            else:
                al = torch.zeros((ny, nc), dtype=torch.float64, device=dev)
            alpha = al[None, :, :].expand(nx, ny, nc).contiguous()
        res = (P, out2, alpha, dxx, dyy)
        if to_host:
            res = _outs(res)
    return res


def pair_stats(K, n):
    """single=4 with dy=None (association.py:517-556 for ``prody is None``): every pair x < y is tested
    with ALL other rows and the covariates held fixed.  With K = the inverse of the Gram matrix of the
    covariate-residualised rows (the dx block of the precision matrix of [dx; dc]), the conditional 2 x 2
    Gram matrix of (x, y) given the rest is the inverse of [[K_xx, K_xy], [K_xy, K_yy]], so
        dxx = K_yy / det / n,  dyy = K_xx / det / n,  dxy = -K_xy / det / n,  det = K_xx K_yy - K_xy^2,
        gamma = dxy / dxx = -K_xy / K_yy,   R^2 = dxy^2 / (dxx dyy) = K_xy^2 / (K_xx K_yy)
    instead of one pseudo-inverse of an (nx + nc - 2) matrix per pair.  Returns (r2, gamma, dyy), (nx, nx),
    meaningful for x != y."""
    kd = torch.diagonal(K)
    kk = kd[:, None] * kd[None, :]
    det = kk - K * K
    return K * K / kk, -K / kd[None, :], kd[:, None] / det / n


def pair_alpha(K, B):
    """alpha of the same test when lowmem is off (association.py:551-553): coefficients on the covariates of
    the regressions of y and of x on everything but the pair, alpha = c_y - gamma c_x.  B (nx, nc) = the
    regression coefficients of every row on the covariates alone; the covariate columns of the precision
    matrix are -K B, and leaving one variable out of a regression is a rank-one update of it."""
    kd = torch.diagonal(K)
    Oc = -(K @ B)                                                    # (nx, nc)
    ratio_y = K / kd[None, :]                                        # K_xy / K_yy
    ratio_x = K / kd[:, None]                                        # K_xy / K_xx
    cx = -(Oc[:, None, :] - ratio_y[:, :, None] * Oc[None, :, :]) / (kd[:, None] - K * ratio_y)[:, :, None]
    cy = -(Oc[None, :, :] - ratio_x[:, :, None] * Oc[:, None, :]) / (kd[None, :] - K * ratio_x)[:, :, None]
    gamma = -ratio_y
    return cy - gamma[:, :, None] * cx


def association_tests_single4_same(dx, dc, lowmem=True, return_dot=True, dimreduce=0, device=None, **ka):
    """``association_tests(dx, None, dc, single=4)``: returns (P, dot|gamma, alpha|None, None, vary) with the
    reference's assembly (association.py:1036-1065; note :1040 - the coefficient is multiplied by vary)."""
    from .association import covariate_basis_device, _to_device_f64, _is_dev, _outs
    tol = ka.pop('tol', 1e-8)
    _exact_inverse_only(ka, max(dx.shape[0] - 2, 0) + dc.shape[0])
    for k in ('precision', 'engine', 'exact_groupings'):
        ka.pop(k, None)
    if ka:
        raise TypeError("association_test_4() got an unexpected keyword argument '{}'".format(next(iter(ka))))
    to_host = not _is_dev(dx)
    ctx = engine.context(device if device is not None else (dx.device if _is_dev(dx) else None))
    nx, n = dx.shape
    nc = dc.shape[0]
    if nx == 0 or n == 0:
        raise ValueError('Dimensions in na==0 detected.')
    if nc == 0:
        logging.warning('No covariate dc input.')
    Qt_dev, rank_c, W = covariate_basis_device(ctx, dc, tol=tol)
    if rank_c > MAX_RANK:
        raise NotImplementedError('covariate rank {} > {}'.format(rank_c, MAX_RANK))
    dof = n - 1 - (max(nx - 2, 0) + rank_c) - dimreduce
    if dof <= 0:
        raise RuntimeError('Insufficient number of cells: must be greater than degrees of '
                           'freedom removed + covariate + 1.')
    with torch.cuda.device(ctx.device):
        dev = ctx.device
        xd = _to_device_f64(ctx, dx)
        if rank_c:
            cf, _ = engine.project_coef(ctx, xd, Qt_dev)
        else:
            cf = torch.zeros((nx, 0), dtype=torch.float64, device=dev)
        G = engine.gram_f64(ctx, xd, cf)                             # Gram matrix of the residualised rows
        eye = torch.eye(nx, dtype=torch.float64, device=dev)
        K, status = engine.de4_solve(ctx, G, eye, torch.ones(nx, dtype=torch.float64, device=dev), n, rank_c, 0, tol,
                                     False)[4:6]                     # w = G^-1 I
        if int(status.item()) & 1:
            raise NotImplementedError('single=4 with dy=None needs rows that are linearly independent given the '
                                      'covariates (the rank-deficient case is not accelerated).')
        K = 0.5 * (K + K.T)
        r2, gamma, dyy = pair_stats(K, n)
        off = ~torch.eye(nx, dtype=torch.bool, device=dev)
        if not bool(((r2[off] >= 0) & (r2[off] <= 1 + 1e-8)).all()):                     # :557
            raise AssertionError('R^2 outside [0, 1]: collinear rows?')
        P = engine.pvalue(ctx, torch.where(off, r2, torch.zeros_like(r2)).contiguous(),
                          torch.full((nx,), dof / 2, dtype=torch.float64, device=dev))
        up = torch.triu(torch.ones((nx, nx), dtype=torch.bool, device=dev), 1)
        zero = torch.zeros_like(P)

        def sym(t):                                                  # triu(., 1) + transpose (:1049-1056)
            t = torch.where(up, t, zero)
            return t + t.T
        P = sym(P)
        vary = sym(dyy)
        vary.diagonal().fill_(1.0)                                   # :1054
        out2 = sym(gamma * dyy)                                      # :1040, :1055-1056
        if not return_dot:
            out2 = out2 / vary                                       # :1061
        alpha = None
        if not lowmem:                                               # :1063-1065
            if rank_c:
                al = pair_alpha(K, cf @ torch.from_numpy(W).to(dev))
            else:
                al = torch.zeros((nx, nx, nc), dtype=torch.float64, device=dev)
            al = torch.where(up[:, :, None], al, torch.zeros_like(al))
            alpha = (al + al.transpose(0, 1)).contiguous()
        if not bool(torch.isfinite(P).all() & torch.isfinite(out2).all() & torch.isfinite(vary).all()):
            raise AssertionError('non-finite result')                # :1077
        res = (P, out2, alpha, None, vary)
        if to_host:
            res = _outs(res)
    return res


def loo_stats(Gxx, Gxy, yy, n, rank_c, tol=1e-8):
    """Leave-one-out regression statistics of association_test_4 (association.py:521-544)
    from Gram matrices of covariate-residualised rows: Gxx = Rx Rx^T (nx, nx),
    Gxy = Rx Ry^T (nx, ny), yy = rowsum(Ry^2).  Returns (dxx, dxy, dyy, rank, w) where w are
    the full multiple-regression coefficients.  float64 torch tensors on any device."""
    from .association import inv_rank
    nx = Gxx.shape[0]
    ev = torch.linalg.eigvalsh(Gxx)
    full_rank = bool(ev[0] > tol * ev[-1]) if nx > 1 else bool(ev[0] > 0)
    if not full_rank:
        return _loo_pinv(Gxx, Gxy, yy, n, rank_c, tol, inv_rank)
    K = torch.linalg.inv(Gxx)
    K = 0.5 * (K + K.T)
    w = K @ Gxy                                   # full-regression coefficients (nx, ny)
    kd = torch.diagonal(K)
    dxx = 1.0 / (n * kd)                          # association.py:539-540 (Schur complement)
    dxy = w / kd[:, None] / n                     # :543-544
    qf = (Gxy * w).sum(dim=0)
    dyy = (yy[None, :] - qf[None, :] + w * w / kd[:, None]) / n      # :541-542
    rank = torch.full((nx,), float(nx - 1 + rank_c), dtype=torch.float64, device=Gxx.device)
    return dxx, dxy, dyy, rank, w


def _row_sumsq(R):
    """sum_k res^2 per row from the stored variance (var = mean, with exact zeros mapped to 1)."""
    return R.var * R.n


def _loo_pinv(Gxx, Gxy, yy, n, rank_c, tol, inv_rank=None):
    """Rank-deficient groupings: the reference's per-x pseudo-inverse (association.py:521-544)
    on the residualised Gram matrices, in float64 on the host."""
    if inv_rank is None:
        from .association import inv_rank
    G = Gxx.cpu().numpy()
    Gy = Gxy.cpu().numpy()
    y2 = yy.cpu().numpy()
    nx, ny = Gy.shape
    dxx = np.zeros(nx)
    dxy = np.zeros((nx, ny))
    dyy = np.zeros((nx, ny))
    rank = np.zeros(nx)
    for x in range(nx):
        t0 = [k for k in range(nx) if k != x]
        r = 0
        if t0:
            gi, r = inv_rank(G[np.ix_(t0, t0)], tol=tol)
        if r == 0:
            dxx[x], dyy[x], dxy[x] = G[x, x] / n, y2 / n, Gy[x] / n
        else:
            cx = G[x, t0] @ gi
            dxx[x] = (G[x, x] - cx @ G[t0, x]) / n
            cy = Gy[t0].T @ gi
            dyy[x] = (y2 - np.sum(cy.T * Gy[t0], axis=0)) / n
            dxy[x] = (Gy[x] - cy @ G[t0, x]) / n
        rank[x] = r + rank_c
    dev = Gxx.device
    w = torch.from_numpy(np.linalg.pinv(G, rcond=tol) @ Gy).to(dev)
    return (torch.from_numpy(dxx).to(dev), torch.from_numpy(dxy).to(dev), torch.from_numpy(dyy).to(dev),
            torch.from_numpy(rank).to(dev), w)
