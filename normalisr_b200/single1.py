"""single=1 (low-MOI screens): every grouping x is tested on its own subset of cells.

Reference: ``association_test_2`` (src/normalisr/association.py:263-390) driven from
``association_tests`` (:910-925): S_x = {cells where dx[x] equals the column sum of dx} = U + T_x
with U the cells without any gRNA (shared by all x) and T_x the cells carrying ONLY x (disjoint
between x).  For every x the reference copies dy[:, S_x], computes a pseudo-inverse of the
covariate Gram matrix of S_x and projects the covariates out of x and of every y again.

Here nothing is copied per x.  Every statistic of that test is a sum over S_x of products of
covariates, x and y, and sum_{S_x} = sum_U + sum_{T_x}:
  * one dense masked pass over dy gives the U parts for all genes (``nsr_project_coef`` with the
    covariates masked to U, FP64 tensor cores);
  * the columns of dy outside U are gathered once in group order and ``nsr_group_stats`` gives the
    T_x parts for all x in one pass over them;
  * per x: pseudo-inverse of its nc x nc Gram matrix with the reference's rank rule
    (``nsr_sym_pinv``, batched Jacobi), then gamma, R^2, d.o.f. and the exact P-value in closed
    form for all genes at once (``nsr_single1_finish``).
With C = sum dc dc^T, cx = sum dc x, cy = sum dc y over S_x and C+ the pseudo-inverse:
  Sxx = sum x^2 - cx^T C+ cx,  Syy = sum y^2 - cy^T C+ cy,  Sxy = sum x y - cx^T C+ cy,
  var_x = Sxx / ns, var_y = Syy / ns, gamma = Sxy / (ns var_x), R^2 = gamma^2 var_x / var_y,
  alpha = C+ cy - gamma C+ cx            (association.py:352-368).
"""
import logging

import numpy as np
import torch

from . import engine
from ._lib import MAX_RANK

_CHUNK_BYTES = 1 << 29        # bound on the (groupings, genes, covariates) temporaries per gene chunk


def _pinv_rank_batched(m, tol=1e-8):
    """inv_rank (association.py:66-80) for a stack of symmetric matrices: singular values below
    tol * largest are dropped; returns (pseudo-inverses, ranks)."""
    u, sv, vt = np.linalg.svd(m)
    keep = sv >= tol * sv[:, :1]
    keep &= sv > 0
    inv = np.where(keep, 1.0 / np.where(keep, sv, 1.0), 0.0)
    return np.einsum('xji,xj,xjk->xik', vt, inv, vt), keep.sum(axis=1).astype(np.float64)


def association_tests_single1(dx, dy, dc, lowmem=True, return_dot=True, dimreduce=0, device=None, **ka):
    from .association import inv_rank, _as_host_f64, _is_dev, _outs
    if ka:
        raise TypeError("association_test_2() got an unexpected keyword argument '{}'".format(next(iter(ka))))
    to_host = not _is_dev(dy)
    ctx = engine.context(device if device is not None else (dy.device if _is_dev(dy) else None))
    dev = ctx.device
    nx, n = dx.shape
    ny, nc = dy.shape[0], dc.shape[0]
    if nc == 0:
        logging.warning('No covariate dc input.')
    if nc + 1 > MAX_RANK:
        raise NotImplementedError('single=1 handles up to {} covariates'.format(MAX_RANK - 1))

    with torch.cuda.device(dev):
        # ---- design: groups of cells                               (association.py:913-918)
        dx_d = (dx if _is_dev(dx) else torch.from_numpy(np.ascontiguousarray(_as_host_f64(dx)))).to(dev, torch.float64)
        dc_d = (dc if _is_dev(dc) else torch.from_numpy(np.ascontiguousarray(_as_host_f64(dc)))).to(dev, torch.float64)
        assert float(dx_d.max()) == 1
        if not bool(((dx_d == 0) | (dx_d == 1)).all()):
            raise NotImplementedError('single=1 is accelerated for binary (0/1) groupings only.')
        colsum = dx_d.sum(dim=0)
        in_u = colsum == 0
        owner = torch.where(colsum == 1, dx_d.argmax(dim=0), torch.full_like(colsum, nx, dtype=torch.int64))
        non_u = torch.nonzero(~in_u).squeeze(1)
        t_order = non_u[torch.sort(owner[non_u], stable=True).indices]           # non-U cells in group order
        counts = torch.bincount(owner[non_u], minlength=nx + 1)
        t_goff = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts, 0)])
        n_t = counts[:nx].to(torch.float64)
        n_u = in_u.sum().to(torch.float64)
        t_ns = n_u + n_t

        # ---- per-x covariate algebra, all on the device: Gram matrices of S_x = U + T_x, their
        # pseudo-inverses with the reference's rank rule (nsr_sym_pinv)   (association.py:343-357)
        u_row = in_u.to(torch.float64)
        ones = torch.ones((1, n), dtype=torch.float64, device=dev)
        q_u = torch.cat([dc_d * u_row, u_row[None]], 0).contiguous()             # covariates masked to U (+ mask)
        c_s = torch.cat([dc_d, ones], 0)[:, t_order].contiguous()                # (nc + 1, m) in group order
        if nc:
            gst = engine.group_stats(ctx, c_s[:nc], c_s, t_goff)[:nx]            # (nx, nc, nc + 2)
            gram = engine.cov_gram(ctx, q_u[:nc])[None] + gst[:, :, :nc]
            t_cx = gst[:, :, nc].contiguous()
            if nc <= 16:
                t_ci, rank = engine.sym_pinv(ctx, gram)
                rank = rank.to(torch.float64)
            else:
                ci_h, rank_h = _pinv_rank_batched(gram.cpu().numpy())
                t_ci, rank = torch.from_numpy(ci_h).to(dev), torch.from_numpy(rank_h).to(dev)
        else:
            t_ci = torch.zeros((nx, 0, 0), dtype=torch.float64, device=dev)
            t_cx = torch.zeros((nx, 0), dtype=torch.float64, device=dev)
            rank = torch.zeros(nx, dtype=torch.float64, device=dev)
        t_ccx = torch.einsum('xij,xj->xi', t_ci, t_cx).contiguous()              # C+ cx
        t_vx = (n_t - (t_cx * t_ccx).sum(dim=1)) / t_ns
        t_vx = torch.where(t_vx == 0, torch.ones_like(t_vx), t_vx)               # :359-361
        t_dof = (t_ns - 1 - rank - dimreduce).contiguous()
        # one synchronisation for the design checks                       (association.py:917-918, 373-376)
        chk = torch.stack([(n_u > 0).to(torch.float64), (n_t > 0).all().to(torch.float64),
                           (t_dof > 0).all().to(torch.float64)]).cpu().numpy()
        assert chk[0] and chk[1]                                  # both values of x occur in S_x
        if not chk[2]:
            raise RuntimeError('Insufficient number of cells: must be greater than degrees of freedom '
                               'removed + covariate + 1.')
        has_rank = (rank > 0)[:, None, None]
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        dof = t_dof

        P = torch.empty((nx, ny), dtype=torch.float64, device=dev)
        gamma = torch.empty_like(P)
        vy = torch.empty_like(P)
        alpha = None if lowmem else torch.zeros((nx, ny, nc), dtype=torch.float64, device=dev)
        step = max(1, min(ny, _CHUNK_BYTES // (8 * nx * (nc + 2))))
        y_host = None
        if not _is_dev(dy):
            y_host = dy if isinstance(dy, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dy))
        for g0 in range(0, ny, step):
            g1 = min(ny, g0 + step)
            if y_host is not None:
                y = y_host[g0:g1].to(dev, torch.float64, non_blocking=True)
            else:
                y = dy[g0:g1].to(torch.float64)
                if y.stride(1) != 1:
                    y = y.contiguous()
            coef_u, yy_all = engine.project_coef(ctx, y, q_u)                    # (gc, nc + 1), (gc,)
            ys = y.index_select(1, t_order)                                      # non-U columns in group order
            st = engine.group_stats(ctx, ys, c_s, t_goff)                        # (nx + 1, gc, nc + 2)
            yy_u = yy_all - st[:, :, nc + 1].sum(dim=0)
            if nc <= 16:                                                         # fused closed form + P-value
                engine.single1_finish(ctx, coef_u, yy_u, st, nx, nc, t_ci, t_cx, t_ccx, t_ns, t_vx, t_dof,
                                      P, gamma, vy, alpha, g0, flag)
                continue
            cy = coef_u[None, :, :nc] + st[:nx, :, :nc]                          # (nx, gc, nc)
            xy = st[:nx, :, nc]                                                  # sum_{T_x} y
            yy = yy_u[None, :] + st[:nx, :, nc + 1]
            ccy = torch.matmul(cy, t_ci)                                         # C+ is symmetric
            syy = yy - (ccy * cy).sum(dim=-1)
            sxy = xy - torch.einsum('xgc,xc->xg', ccy, t_cx)
            v_y = syy / t_ns[:, None]
            gam = sxy / (t_ns * t_vx)[:, None]                                   # :364
            r2 = gam * gam * t_vx[:, None] / v_y                                 # :368
            if not bool(((r2 >= 0) & (r2 <= 1 + 1e-8)).all()):                   # :371
                raise AssertionError('R^2 outside [0, 1].')
            P[:, g0:g1] = engine.pvalue(ctx, r2.clamp(max=1.0), dof / 2)
            gamma[:, g0:g1] = gam
            vy[:, g0:g1] = v_y
            if alpha is not None:                                                # :365-367
                al = ccy - gam[:, :, None] * t_ccx[:, None, :]
                alpha[:, g0:g1] = torch.where(has_rank, al, torch.zeros_like(al))
        if int(flag.item()):
            raise AssertionError('R^2 outside [0, 1].')                          # association.py:371
        out2 = gamma * t_vx[:, None] if return_dot else gamma                    # association.py:1058-1061
        res = (P, out2, alpha, t_vx, vy)
        if to_host:
            res = _outs(res)
    return res
