"""``normalisr.norm.normvar`` on the GPU (reference src/normalisr/norm.py:131-289), the step
directly upstream of ``coex`` / ``de``: mean and variance normalisation of the
log-CPM matrix.  Gene x is multiplied by ``w ** wt[x]`` per cell and its OWN weighted covariates
``dc * w ** wt[x]`` are projected out (one pseudo-inverse per gene, ``normvar1``); optionally the
variance of every gene is restored (``keepvar``), continuous covariates are scaled by ``w``.

The reference loops over genes in Python.  Here the per-gene Gram matrices and right-hand sides
come from one streaming pass over dt (``nsr_normvar_stats``: two skinny GEMMs over cells on the
FP64 tensor cores, because G_x = sum_k s^2 (c c^T) depends on the gene only through s), the nc x nc pseudo-inverses are
batched on the device (``nsr_sym_pinv``), and a second pass writes the result (``nsr_normvar_apply``); the residual
variance needed by ``keepvar`` follows from the same statistics (S2 - b^T G+ b), so there is no
third pass.  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.
"""
import numpy as np
import torch

from . import _lib, engine

_ROW_CHUNK_BYTES = 1 << 30


def _is_dev(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def _dev64(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev, torch.float64)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev, torch.float64)


def _design(ctx, dc_d):
    """Gene-independent right-hand matrix of the statistics GEMM: the nc (nc + 1) / 2 products
    c_i c_j (row-major upper triangle), zero rows up to a multiple of 8 tiles, then the covariates
    themselves, zero rows up to 16.  Returns (M, tri, first row of the covariate part)."""
    nc, n = dc_d.shape
    cols = ctx.lib.nsr_normvar_width(nc)                  # 8 * (D tiles + 2) + 2
    d_rows = cols - 2 - 16
    iu = torch.triu_indices(nc, nc, device=dc_d.device)
    M = torch.zeros((d_rows + 16, n), dtype=torch.float64, device=dc_d.device)
    M[:iu.shape[1]] = dc_d[iu[0]] * dc_d[iu[1]]
    M[d_rows:d_rows + nc] = dc_d
    return M, iu, d_rows


def _normvar_rows(ctx, dt_d, dc_d, design, logw, wt_d, keepvar):
    """One block of genes resident on the device -> normalised block."""
    genes, n = dt_d.shape
    nc = dc_d.shape[0]
    M, iu, d_rows = design
    tri = iu.shape[1]
    stats = torch.empty((genes, M.shape[0] + 2), dtype=torch.float64, device=dt_d.device)
    ld = dt_d.stride(0) if genes > 1 else n
    ldc = dc_d.stride(0) if nc > 1 else n
    _lib.check(ctx.lib.nsr_normvar_stats(ctx.handle, engine._stream(), dt_d.data_ptr(), genes, n, ld, M.data_ptr(), nc,
                                         M.stride(0), logw.data_ptr(), wt_d.data_ptr(), stats.data_ptr()),
               "nsr_normvar_stats")
    G = torch.zeros((genes, nc, nc), dtype=torch.float64, device=dt_d.device)
    G[:, iu[0], iu[1]] = stats[:, :tri]
    G = G + torch.triu(G, 1).transpose(1, 2)
    b = stats[:, d_rows:d_rows + nc]
    s1, s2 = stats[:, M.shape[0]], stats[:, M.shape[0] + 1]
    ci, rank = engine.sym_pinv(ctx, G)                                      # inv_rank per gene, norm.py:159-160
    if bool((rank <= 0).any()):
        raise RuntimeError('Zero-rank covariates found.')                    # norm.py:161-162
    coef = torch.einsum('gij,gj->gi', ci, b).contiguous()
    if keepvar:                                                              # norm.py:241-243, 251-254
        dv = torch.sqrt(s2 / n - (s1 / n) ** 2)
        dv2 = torch.sqrt((s2 - (b * coef).sum(dim=1)) / n)
        scale = (dv / dv2) ** wt_d
    else:
        scale = torch.ones(genes, dtype=torch.float64, device=dt_d.device)
    # the reference asserts that its output is finite (norm.py:277); S2 = sum (s dt)^2, coef and scale
    # finite imply it, without another pass over the matrix
    if not bool(torch.isfinite(stats).all() & torch.isfinite(coef).all() & torch.isfinite(scale).all()):
        raise AssertionError('non-finite values in the normalised expression matrix')
    out = torch.empty((genes, n), dtype=torch.float64, device=dt_d.device)
    _lib.check(ctx.lib.nsr_normvar_apply(ctx.handle, engine._stream(), dt_d.data_ptr(), genes, n, ld, dc_d.data_ptr(), nc,
                                         ldc, logw.data_ptr(), wt_d.data_ptr(), coef.data_ptr(), scale.contiguous().data_ptr(),
                                         out.data_ptr(), n), "nsr_normvar_apply")
    engine.LAUNCHES += 3
    return out


def normvar(dt, dc, w, wt, dextra=None, cat=1, nth=1, bs=500, keepvar=True, normmean=False, device=None):
    """Performs mean and variance normalisations; same arguments, checks and return value as the
    reference (norm.py:169-289): ``[dtn, dcn]`` or ``[dtn, dcn, dextran]``.  ``nth`` / ``bs`` are
    accepted and ignored."""
    from .association import inv_rank
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
    if float(w.min()) <= 0:
        raise ValueError('w must be positive.')
    if float(wt.min()) < 0:
        raise ValueError('wt must be non-negative.')
    if cat not in (0, 1, 2):
        raise ValueError('Invalid cat value.')
    nc = dc.shape[0]
    if nc > 12:
        raise NotImplementedError('normvar is accelerated for up to 12 covariates.')
    to_host = not _is_dev(dt)
    ctx = engine.context(device if device is not None else (dt.device if _is_dev(dt) else None))
    dev = ctx.device
    with torch.cuda.device(dev):
        dc_d = _dev64(dc, dev).contiguous()
        w_d = _dev64(w, dev).contiguous()
        wt_d = _dev64(wt, dev).contiguous()
        logw = torch.log(w_d)
        design = _design(ctx, dc_d)
        # covariates: continuous rows (and, for cat = 1, the intercept) are scaled by w   norm.py:257-269
        if cat == 2:
            sel = torch.ones(nc, dtype=torch.bool, device=dev)
        else:
            sel = ((dc_d != 0) & (dc_d != 1)).any(dim=1)
            if cat == 1:
                sel = sel | (dc_d == 1).all(dim=1)
        dcn = torch.where(sel[:, None], dc_d * w_d, dc_d)
        # expression, in row blocks (genes are independent)
        if to_host:
            src = dt if isinstance(dt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dt))
            dtn = torch.empty((nt, ns), dtype=torch.float64)
            step = max(1, _ROW_CHUNK_BYTES // (8 * ns))
        else:
            src = dt.to(torch.float64)
            if src.stride(1) != 1:
                src = src.contiguous()
            dtn = torch.empty((nt, ns), dtype=torch.float64, device=dev)
            step = nt
        if normmean:                                                         # norm.py:271-273 (normvar1 on dcn)
            gi, r = inv_rank(engine.cov_gram(ctx, dcn.contiguous()).cpu().numpy())
            if r <= 0:
                raise RuntimeError('Zero-rank covariates found.')
            gi_d = torch.from_numpy(gi).to(dev)
        for g0 in range(0, nt, step):
            g1 = min(nt, g0 + step)
            blk = src[g0:g1].to(dev, torch.float64, non_blocking=True) if to_host else src[g0:g1]
            res = _normvar_rows(ctx, blk, dc_d, design, logw, wt_d[g0:g1].contiguous(), keepvar)
            if normmean:
                cf, _ = engine.project_coef(ctx, res, dcn.contiguous())
                res.addmm_(cf @ gi_d, dcn, alpha=-1.0)
            dtn[g0:g1] = res if not to_host else res.cpu()
        if not bool(torch.isfinite(dcn).all()):
            raise AssertionError('non-finite values in the normalised covariates')  # norm.py:277
        ans = [dtn, dcn]
        if dextra is not None:
            ans.append(_dev64(dextra, dev) * w_d)
        if to_host:
            ans = [a.cpu().numpy() if a.is_cuda else a.numpy() for a in ans]
    return ans


def compute_var(dt, dc, stepmax=1, eps=1E-6, device=None):
    """Computes the variance normalisation multiplier of every cell (reference norm.py:56-128): a
    log-linear fit of each cell's residual variance on the covariates.  Same arguments, checks and
    return value (``(n_cell,)`` array, minimum 1).  Only ``stepmax=1`` (the reference's default, no
    EM-like iterations) is accelerated.

    The residual matrix is never materialised: one pass over dt gives every gene's projection
    coefficients, residual mean and variance (``nsr_project_coef`` on the orthonormal covariate
    basis plus a row of ones), a second one the per-cell sums of squared standardised residuals
    (``nsr_colvar``); the fit of the log variances on the covariates is a rank-sized problem."""
    from .association import covariate_basis_device
    if eps <= 0 or stepmax <= 0:
        raise ValueError('eps and stepmax must be positive.')
    if dt.ndim != 2 or dc.ndim != 2:
        raise ValueError('dt and dc must both have 2 dimensions.')
    if dt.shape[1] != dc.shape[1]:
        raise ValueError('dt and dc must have the same cell count.')
    if stepmax != 1:
        raise NotImplementedError('compute_var is accelerated for stepmax=1 (no EM-like iterations).')
    to_host = not _is_dev(dt)
    ctx = engine.context(device if device is not None else (dt.device if _is_dev(dt) else None))
    dev = ctx.device
    nt, ns = dt.shape
    with torch.cuda.device(dev):
        dc_d = _dev64(dc, dev).contiguous()
        Qt, rank, _ = covariate_basis_device(ctx, dc_d)           # least squares without intercept = projection
        if rank > 16:
            raise NotImplementedError('compute_var is accelerated for covariate rank <= 16.')
        ones = torch.ones((1, ns), dtype=torch.float64, device=dev)
        q1 = torch.cat([Qt, ones], 0).contiguous() if rank else ones
        qsum = Qt.sum(dim=1) if rank else None
        src = dt if not to_host else (dt if isinstance(dt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dt)))
        step = nt if not to_host else max(1, _ROW_CHUNK_BYTES // (8 * ns))
        col = torch.zeros(ns, dtype=torch.float64, device=dev)
        for g0 in range(0, nt, step):
            g1 = min(nt, g0 + step)
            blk = src[g0:g1].to(dev, torch.float64, non_blocking=True)
            if blk.stride(1) != 1:
                blk = blk.contiguous()
            cf, sxx = engine.project_coef(ctx, blk, q1)                       # (gc, rank + 1), (gc,)
            coef = cf[:, :rank]
            mean = (cf[:, rank] - (coef @ qsum if rank else 0.0)) / ns        # norm.py:101
            var = (sxx - (coef * coef).sum(dim=1)) / ns - mean * mean         # :102 (Qt is orthonormal)
            istd = (1.0 / torch.sqrt(var)).contiguous()
            part = torch.empty(ns, dtype=torch.float64, device=dev)
            _lib.check(ctx.lib.nsr_colvar(ctx.handle, engine._stream(), blk.data_ptr(), g1 - g0, ns,
                                          blk.stride(0) if g1 - g0 > 1 else ns, Qt.data_ptr() if rank else None, rank,
                                          Qt.stride(0) if rank > 1 else ns, cf.data_ptr(), cf.stride(0),
                                          mean.contiguous().data_ptr(), istd.data_ptr(), part.data_ptr()), "nsr_colvar")
            engine.LAUNCHES += 2
            col += part
        y = torch.log(torch.sqrt(col / nt))                                   # :103
        # log-linear fit with intercept (:104-105): mean + projection on the centred covariates
        xc = dc_d - dc_d.mean(dim=1, keepdim=True)
        Qc, rc, _ = covariate_basis_device(ctx, xc)
        ym = y.mean()
        pred = ym + (Qc.T @ (Qc @ (y - ym)) if rc else 0.0)
        new = torch.exp(pred)                                                 # :107 (scale was 1)
        new = new / new.min()
        w = 1.0 / new                                                         # :122-123
        w = w / w.min()
        if not bool(torch.isfinite(w).all() & (w > 0).all()):
            raise AssertionError('non-finite variance normalisation multipliers')
        return w.cpu().numpy() if to_host else w
