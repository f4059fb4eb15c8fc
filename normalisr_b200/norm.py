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

from . import _lib, engine, hoststage

_ROW_CHUNK_BYTES = 1 << 30
_CENTRED = {}              # device index -> (key, (Qc, rank), tensor): basis of the centred covariates (compute_var)
_USE_CHEBYSHEV = True      # test hook: False = Gram matrices from the streaming pass (nsr_normvar_stats)


_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device, priority=-1)
    return _SIDE[key]


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


_CHEB_NODES = 24           # Chebyshev nodes per piece of the wt range
_CHEB_ALPHA = 1.5          # largest half-range of the exponent within a piece
_CHEB_PIECES = 256         # beyond this the streaming pass computes the Gram matrices


def _cheb_pieces(bounds):
    """Number of pieces the wt range is cut into for (min w, max w, min wt, max wt); 0 = not applicable."""
    if not all(np.isfinite(v) for v in bounds) or bounds[0] <= 0:
        return 0
    lw_half = 0.5 * (np.log(bounds[1]) - np.log(bounds[0])) * (1 + 1e-12)
    pieces = max(1, int(np.ceil(0.5 * (bounds[3] - bounds[2]) * 2.0 * lw_half / _CHEB_ALPHA)))
    return pieces if pieces <= _CHEB_PIECES else 0


def _gram_chebyshev(dc_d, logw, wt_d, bounds=None):
    """Per-gene covariate Gram matrices G_x = sum_k w_k ** (2 wt_x) c_k c_k^T (norm.py:156-159: dc * w2[x]
    times its transpose) for ALL genes without a pass over the expression matrix.

    G depends on the gene only through the scalar wt_x:  G(wt) = e^(2 wt m) sum_k e^(2 wt (lw_k - m)) c_k c_k^T
    with lw = log w and m the mid-range of lw.  The sum is an entire function of wt; on an interval of
    half-width h it is a Chebyshev series in x = (wt - centre) / h whose coefficients decay like
    I_j(alpha), alpha = h (max lw - min lw).  So: cut the range of wt into pieces with alpha <= 1.5, evaluate
    the sum at 24 Chebyshev nodes per piece (one (24 x n) x (n x nc (nc + 1) / 2) product per piece;
    I_24(1.5) / I_0(1.5) ~ 1e-27), and interpolate every gene with the barycentric formula (backward
    stable; Higham 2004).  Interpolation is accurate relative to the largest value on the piece, and the
    terms of the sum grow at rates between -alpha and +alpha across it, so the pointwise relative error is
    about e^(2 alpha) roundings: hence the small alpha (measured: 1e-15 of sum |terms|, like a direct sum).
    Returns (G, ok): ok is a device flag - the truncated Chebyshev coefficients have decayed to rounding
    level (read later by the caller, so that nothing here waits for the device) - or None when the weights
    span too wide a range (the caller then takes the streaming-pass statistics).  ``bounds`` = (min w,
    max w, min wt, max wt) if the caller already has them on the host."""
    dev = dc_d.device
    nc, n = dc_d.shape
    genes = wt_d.shape[0]
    iu = torch.triu_indices(nc, nc, device=dev)
    D = (dc_d[iu[0]] * dc_d[iu[1]]).t().contiguous()                # (n, tri)
    if bounds is None:
        bounds = torch.stack([torch.exp(logw.min()), torch.exp(logw.max()), wt_d.min(), wt_d.max()]).cpu().tolist()
    pieces = _cheb_pieces(bounds)
    if not pieces:
        return None
    lo, hi = float(np.log(bounds[0])), float(np.log(bounds[1]))
    mid = 0.5 * (lo + hi)
    wmin, wmax = float(bounds[2]), float(bounds[3])
    h = 0.5 * (wmax - wmin) / pieces
    m = _CHEB_NODES
    j = torch.arange(m, dtype=torch.float64, device=dev)
    xn = torch.cos(np.pi * (j + 0.5) / m)                           # nodes (first kind)
    bw = torch.sin(np.pi * (j + 0.5) / m) * (1.0 - 2.0 * (j % 2))    # barycentric weights
    centres = wmin + h * (2.0 * torch.arange(pieces, dtype=torch.float64, device=dev) + 1.0)
    piece = torch.clamp(((wt_d - wmin) / (2.0 * h)).floor().long(), 0, pieces - 1) if h > 0 else \
        torch.zeros(genes, dtype=torch.long, device=dev)
    x = ((wt_d - centres[piece]) / h).clamp_(-1.0, 1.0) if h > 0 else torch.zeros_like(wt_d)
    lw0 = logw - mid
    wt_nodes = (centres[:, None] + h * xn[None, :]).reshape(-1)     # (pieces * m,)
    Dabs = D.abs()
    F = torch.empty((pieces * m, D.shape[1]), dtype=torch.float64, device=dev)      # node values
    Fabs = torch.empty_like(F)                                        # their rounding scale (sum of |terms|)
    step = max(m, (1 << 25) // max(1, n) // m * m)                    # rows of E per batch (256 MB)
    for r0 in range(0, pieces * m, step):
        E = torch.exp(2.0 * wt_nodes[r0:r0 + step, None] * lw0[None, :])
        F[r0:r0 + step] = E @ D
        Fabs[r0:r0 + step] = E @ Dabs
    del E
    F, Fabs = F.reshape(pieces, m, -1), Fabs.reshape(pieces, m, -1)
    # decay of the Chebyshev coefficients (type-II DCT of the node values): the last four must be at rounding level
    i = torch.arange(m, dtype=torch.float64, device=dev)
    A = (2.0 / m) * torch.einsum('ij,pjt->pit', torch.cos(np.pi * i[:, None] * (j[None, :] + 0.5) / m), F)
    scale = Fabs.amax(dim=1)                                         # (pieces, tri)
    ok = (A[:, -4:].abs().amax(dim=1) <= 1e-14 * scale).all()
    diff = x[:, None] - xn[None, :]                                  # (genes, m)
    hit = diff == 0
    wgt = bw[None, :] / torch.where(hit, torch.ones_like(diff), diff)
    any_hit = hit.any(dim=1, keepdim=True)
    wgt = torch.where(any_hit, hit.to(torch.float64), wgt)           # a gene exactly on a node takes the node value
    wgt = wgt / wgt.sum(dim=1, keepdim=True)
    tri = torch.empty((genes, D.shape[1]), dtype=torch.float64, device=dev)
    if pieces == 1:
        tri = wgt @ F[0]
    else:
        for p in range(pieces):
            idx = (piece == p).nonzero(as_tuple=True)[0]
            if idx.numel():
                tri[idx] = wgt[idx] @ F[p]
    tri = tri * torch.exp(2.0 * mid * wt_d)[:, None]
    G = torch.zeros((genes, nc, nc), dtype=torch.float64, device=dev)
    G[:, iu[0], iu[1]] = tri
    return G + torch.triu(G, 1).transpose(1, 2), ok


_WIDE_BYTES = 1 << 28      # bound on each (genes, cells) temporary of the wide fallbacks below


def _pinv_rank_sym(G, tol=1e-8):
    """inv_rank (association.py:66-80) for a stack of symmetric matrices on the device: the singular values of a
    symmetric matrix are the absolute eigenvalues and its right singular vectors the eigenvectors, so
    ``(Vt.T / s) @ Vt`` with values below tol * largest dropped is V diag(1 / |lambda|) V^T over the kept ones."""
    lam, V = torch.linalg.eigh(G)
    a = lam.abs()
    keep = (a >= tol * a.amax(dim=-1, keepdim=True)) & (a > 0)
    inv = torch.where(keep, 1.0 / torch.where(keep, a, torch.ones_like(a)), torch.zeros_like(a))
    return (V * inv[..., None, :]) @ V.transpose(-1, -2), keep.sum(dim=-1)


def _normvar_rows_wide(dt_d, dc_d, logw, wt_d, keepvar, G, out, flags):
    """normvar for more covariates than the kernels stage (nc > 16): the same algebra as ``_normvar_rows`` with
    library products (float64 GEMMs, batched symmetric eigendecompositions) in gene chunks.  ``G``: the chunk's
    Gram matrices if the interpolation supplies them, else None (formed here, one GEMM on the products c_i c_j)."""
    genes, n = dt_d.shape
    nc = dc_d.shape[0]
    step = max(1, _WIDE_BYTES // (8 * n))
    D = None
    for g0 in range(0, genes, step):
        g1 = min(genes, g0 + step)
        wt = wt_d[g0:g1]
        s = torch.exp(wt[:, None] * logw[None, :])                          # w ** wt (1 where wt == 0), norm.py:238-239
        xs = dt_d[g0:g1] * s
        b = (xs * s) @ dc_d.T                                               # sum_k (s dt) (s c)
        if G is not None:
            Gc = G[g0:g1]
        else:
            if D is None:
                iu = torch.triu_indices(nc, nc, device=dc_d.device)
                D = (dc_d[iu[0]] * dc_d[iu[1]]).T.contiguous()              # (n, tri)
            tri = (s * s) @ D
            Gc = torch.zeros((g1 - g0, nc, nc), dtype=torch.float64, device=dc_d.device)
            Gc[:, iu[0], iu[1]] = tri
            Gc = Gc + torch.triu(Gc, 1).transpose(1, 2)
        ci, rank = _pinv_rank_sym(Gc)                                        # norm.py:159-160
        coef = torch.einsum('gij,gj->gi', ci, b)
        res = xs - s * (coef @ dc_d)                                         # norm.py:163 with dc * w2[x]
        if keepvar:                                                          # norm.py:241-243, 251-254
            dv = torch.sqrt(((xs - xs.mean(dim=1, keepdim=True)) ** 2).mean(dim=1))
            dv2 = torch.sqrt((res * res).mean(dim=1))
            res = res * ((dv / dv2) ** wt)[:, None]
        flags["zero_rank"].append((rank <= 0).any())
        flags["finite"].append(torch.isfinite(res).all())
        out[g0:g1] = res
    return out


def _normvar_rows(ctx, dt_d, dc_d, design, logw, wt_d, keepvar, G=None, out=None, flags=None):
    """One block of genes resident on the device -> normalised block.  ``G``: a callable returning the
    block's Gram matrices (``_gram_chebyshev``; called AFTER the statistics kernel is queued, so its many
    small launches are prepared while that kernel runs), or None: they come out of the streaming pass as
    well (``nsr_normvar_stats``).  ``flags``: dict of lists that receive device booleans (zero-rank
    covariates, non-finite results) for the caller to read once, instead of a synchronisation each."""
    genes, n = dt_d.shape
    nc = dc_d.shape[0]
    ld = dt_d.stride(0) if genes > 1 else n
    ldc = dc_d.stride(0) if nc > 1 else n
    if G is not None:
        C16 = design
        stats = torch.empty((genes, 18), dtype=torch.float64, device=dt_d.device)
        cur = torch.cuda.current_stream(dt_d.device)
        inputs_ready = torch.cuda.Event()
        inputs_ready.record(cur)
        _lib.check(ctx.lib.nsr_normvar_rhs(ctx.handle, engine._stream(), dt_d.data_ptr(), genes, n, ld, C16.data_ptr(),
                                           C16.stride(0), logw.data_ptr(), wt_d.data_ptr(), stats.data_ptr()),
                   "nsr_normvar_rhs")
        # the Gram matrices do not depend on the statistics kernel: their ~60 small launches run on a high-priority
        # side stream, in the gaps of that kernel instead of behind it
        side = _side_stream(dt_d.device)
        side.wait_event(inputs_ready)
        with torch.cuda.stream(side):
            G = G()
        cur.wait_stream(side)
        G.record_stream(cur)
        b = stats[:, :nc]
        s1, s2 = stats[:, 16], stats[:, 17]
    else:
        M, iu, d_rows = design
        tri = iu.shape[1]
        stats = torch.empty((genes, M.shape[0] + 2), dtype=torch.float64, device=dt_d.device)
        _lib.check(ctx.lib.nsr_normvar_stats(ctx.handle, engine._stream(), dt_d.data_ptr(), genes, n, ld, M.data_ptr(), nc,
                                             M.stride(0), logw.data_ptr(), wt_d.data_ptr(), stats.data_ptr()),
                   "nsr_normvar_stats")
        G = torch.zeros((genes, nc, nc), dtype=torch.float64, device=dt_d.device)
        G[:, iu[0], iu[1]] = stats[:, :tri]
        G = G + torch.triu(G, 1).transpose(1, 2)
        b = stats[:, d_rows:d_rows + nc]
        s1, s2 = stats[:, M.shape[0]], stats[:, M.shape[0] + 1]
    ci, rank = engine.sym_pinv(ctx, G)                                      # inv_rank per gene, norm.py:159-160
    coef = torch.einsum('gij,gj->gi', ci, b).contiguous()
    if keepvar:                                                              # norm.py:241-243, 251-254
        dv = torch.sqrt(s2 / n - (s1 / n) ** 2)
        dv2 = torch.sqrt((s2 - (b * coef).sum(dim=1)) / n)
        scale = (dv / dv2) ** wt_d
    else:
        scale = torch.ones(genes, dtype=torch.float64, device=dt_d.device)
    # zero-rank covariates (norm.py:161-162); the reference asserts that its output is finite (norm.py:277):
    # S2 = sum (s dt)^2, coef and scale finite imply it, without another pass over the matrix
    zero_rank = (rank <= 0).any()
    finite = torch.isfinite(stats).all() & torch.isfinite(coef).all() & torch.isfinite(scale).all()
    if flags is None:
        if bool(zero_rank):
            raise RuntimeError('Zero-rank covariates found.')
        if not bool(finite):
            raise AssertionError('non-finite values in the normalised expression matrix')
    else:
        flags["zero_rank"].append(zero_rank)
        flags["finite"].append(finite)
    if out is None:
        out = torch.empty((genes, n), dtype=torch.float64, device=dt_d.device)
    _lib.check(ctx.lib.nsr_normvar_apply(ctx.handle, engine._stream(), dt_d.data_ptr(), genes, n, ld, dc_d.data_ptr(), nc,
                                         ldc, logw.data_ptr(), wt_d.data_ptr(), coef.data_ptr(), scale.contiguous().data_ptr(),
                                         out.data_ptr(), out.stride(0) if genes > 1 else n), "nsr_normvar_apply")
    engine.LAUNCHES += 3
    return out


def normvar(dt, dc, w, wt, dextra=None, cat=1, nth=1, bs=500, keepvar=True, normmean=False, device=None,
            _chebyshev=None):
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
    def min_max(x):                       # the checks here and the interpolation below need the four bounds
        if _is_dev(x):
            return torch.stack([x.min(), x.max()]).double().cpu().tolist()
        return [float(x.min()), float(x.max())]
    bounds = min_max(w) + min_max(wt)
    if not bounds[0] > 0 and not np.isnan(bounds[0]):
        raise ValueError('w must be positive.')
    if bounds[2] < 0:
        raise ValueError('wt must be non-negative.')
    if cat not in (0, 1, 2):
        raise ValueError('Invalid cat value.')
    nc = dc.shape[0]
    wide = nc > 16                       # more covariates than the kernels stage: library products (_normvar_rows_wide)
    to_host = not _is_dev(dt)
    ctx = engine.context(device if device is not None else (dt.device if _is_dev(dt) else None))
    dev = ctx.device
    with torch.cuda.device(dev):
        dc_d = _dev64(dc, dev).contiguous()
        w_d = _dev64(w, dev).contiguous()
        wt_d = _dev64(wt, dev).contiguous()
        logw = torch.log(w_d)
        # per-gene Gram matrices by interpolation in wt (no pass over dt); the streaming statistics otherwise.
        # The interpolation is queued after the first block's statistics kernel (see _normvar_rows).
        cheb = {"G": None, "ok": None}
        use_cheb = (_USE_CHEBYSHEV if _chebyshev is None else _chebyshev) and _cheb_pieces(bounds) > 0

        def gram_block(g0, g1):
            if cheb["G"] is None:
                cheb["G"], cheb["ok"] = _gram_chebyshev(dc_d, logw, wt_d, bounds)
            return cheb["G"][g0:g1]
        flags = {"zero_rank": [], "finite": []}
        if wide:
            design = None
        elif use_cheb:
            design = torch.zeros((16, ns), dtype=torch.float64, device=dev)
            design[:nc] = dc_d
        elif nc <= 12:
            design = _design(ctx, dc_d)
        else:
            raise NotImplementedError('normvar with more than 12 covariates needs weights whose range the '
                                      'interpolation of the Gram matrices covers.')
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
            if src.dtype != torch.float64:
                src = src.to(torch.float64)
            dtn = torch.empty((nt, ns), dtype=torch.float64)
            step = max(1, min(nt, min(_ROW_CHUNK_BYTES, hoststage.STAGE_BYTES) // (8 * ns)))
            blocks = hoststage.RowBlocks(ctx, src, step, out_row_bytes=8 * ns, tag="normvar")
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
            blk = blocks.fetch(g0, g1) if to_host else src[g0:g1]
            if wide:
                res = _normvar_rows_wide(blk, dc_d, logw, wt_d[g0:g1].contiguous(), keepvar,
                                         gram_block(g0, g1) if use_cheb else None,
                                         torch.empty((g1 - g0, ns), dtype=torch.float64, device=dev) if to_host else dtn[g0:g1],
                                         flags)
            else:
                res = _normvar_rows(ctx, blk, dc_d, design, logw, wt_d[g0:g1].contiguous(), keepvar,
                                    G=(lambda a=g0, b=g1: gram_block(a, b)) if use_cheb else None,
                                    out=None if to_host else dtn[g0:g1], flags=flags)
            if normmean:
                cf, _ = engine.project_coef(ctx, res, dcn.contiguous())
                res.addmm_(cf @ gi_d, dcn, alpha=-1.0)
            if to_host:
                blocks.store(dtn[g0:g1], res)
        if to_host:
            blocks.close()
        # every deferred check in one read
        chk = torch.stack([torch.stack(flags["zero_rank"]).any(), torch.stack(flags["finite"]).all(),
                           torch.isfinite(dcn).all(),
                           cheb["ok"] if cheb["ok"] is not None else torch.ones((), dtype=torch.bool, device=dev)]).cpu().tolist()
        if not chk[3]:
            # the interpolation did not reach rounding level (not observed; the pieces are sized for it): take
            # the Gram matrices from the streaming pass instead
            return normvar(dt, dc, w, wt, dextra=dextra, cat=cat, keepvar=keepvar, normmean=normmean, device=device,
                           _chebyshev=False)
        if chk[0]:
            raise RuntimeError('Zero-rank covariates found.')                        # norm.py:161-162
        if not chk[1]:
            raise AssertionError('non-finite values in the normalised expression matrix')
        if not chk[2]:
            raise AssertionError('non-finite values in the normalised covariates')  # norm.py:277
        ans = [dtn, dcn]
        if dextra is not None:
            ans.append(_dev64(dextra, dev) * w_d)
        if to_host:
            ans = [a.cpu().numpy() if a.is_cuda else a.numpy() for a in ans]
    return ans


def compute_var(dt, dc, stepmax=1, eps=1E-6, device=None):
    """Computes the variance normalisation multiplier of every cell (reference norm.py:56-128): a
    log-linear fit of each cell's residual variance on the covariates, optionally iterated (EM-like,
    ``stepmax`` > 1, :97-121).  Same arguments, checks and return value (``(n_cell,)`` array, minimum 1).

    The residual matrix is never materialised: one pass over dt gives every gene's projection
    coefficients, residual mean and variance (``nsr_project_coef`` on the orthonormal covariate
    basis plus a row of ones), a second one the per-cell sums of squared standardised residuals
    (``nsr_colvar``); the fit of the log variances on the covariates is a rank-sized problem.
    Iterations after the first work on dt / scale and dc / scale (:99-100); the scaled block is formed
    once per iteration (one more read and write of the block), the rest is the same two passes."""
    from .association import basis_cache_key, covariate_basis_device
    if eps <= 0 or stepmax <= 0:
        raise ValueError('eps and stepmax must be positive.')
    if dt.ndim != 2 or dc.ndim != 2:
        raise ValueError('dt and dc must both have 2 dimensions.')
    if dt.shape[1] != dc.shape[1]:
        raise ValueError('dt and dc must have the same cell count.')
    to_host = not _is_dev(dt)
    ctx = engine.context(device if device is not None else (dt.device if _is_dev(dt) else None))
    dev = ctx.device
    nt, ns = dt.shape
    with torch.cuda.device(dev):
        dc_d = _dev64(dc, dev).contiguous()
        src = dt if not to_host else (dt if isinstance(dt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dt)))
        if to_host and src.dtype != torch.float64:
            src = src.to(torch.float64)
        step = nt if not to_host else max(1, min(nt, min(_ROW_CHUNK_BYTES, hoststage.STAGE_BYTES) // (8 * ns)))
        blocks = hoststage.RowBlocks(ctx, src, step, tag="compute_var") if to_host else None
        ones = torch.ones((1, ns), dtype=torch.float64, device=dev)
        # log-linear fit with intercept (:104-105, on the UNSCALED covariates): mean + projection on the centred covariates
        ckey = basis_cache_key(ctx, dc_d, tag='centred')
        if ckey is not None and _CENTRED.get(dev.index, (None,))[0] == ckey:
            Qc, rc = _CENTRED[dev.index][1]
        else:
            xc = dc_d - dc_d.mean(dim=1, keepdim=True)
            Qc, rc, _ = covariate_basis_device(ctx, xc)
            if ckey is not None:
                _CENTRED[dev.index] = (ckey, (Qc, rc), dc_d)
        scale = None                       # d1sscale; None = all ones (first iteration)
        best, bestv, it = None, 1e300, 0
        while it < stepmax and bestv > eps:
            inv = None if scale is None else (1.0 / scale)
            Qt, rank, _ = covariate_basis_device(ctx, dc_d if inv is None else dc_d * inv)   # least squares without intercept = projection
            q1 = torch.cat([Qt, ones], 0).contiguous() if rank else ones
            qsum = Qt.sum(dim=1) if rank else None
            col = torch.zeros(ns, dtype=torch.float64, device=dev)
            for g0 in range(0, nt, step):
                g1 = min(nt, g0 + step)
                blk = blocks.fetch(g0, g1) if to_host else src[g0:g1].to(dev, torch.float64, non_blocking=True)
                if inv is not None:
                    blk = blk * inv                                               # :99
                if blk.stride(1) != 1:
                    blk = blk.contiguous()
                cf, sxx = engine.project_coef(ctx, blk, q1)                       # (gc, rank + 1), (gc,)
                coef = cf[:, :rank]
                mean = (cf[:, rank] - (coef @ qsum if rank else 0.0)) / ns        # norm.py:104
                var = (sxx - (coef * coef).sum(dim=1)) / ns - mean * mean         # :105 (Qt is orthonormal)
                istd = (1.0 / torch.sqrt(var)).contiguous()
                if rank > 16:
                    # more basis rows than nsr_colvar keeps in registers: the residual block in gene chunks (library GEMM)
                    sub = max(1, _WIDE_BYTES // (8 * ns))
                    for h0 in range(0, g1 - g0, sub):
                        h1 = min(g1 - g0, h0 + sub)
                        r = blk[h0:h1] - coef[h0:h1] @ Qt
                        col += (((r - mean[h0:h1, None]) * istd[h0:h1, None]) ** 2).sum(dim=0)
                    continue
                part = torch.empty(ns, dtype=torch.float64, device=dev)
                _lib.check(ctx.lib.nsr_colvar(ctx.handle, engine._stream(), blk.data_ptr(), g1 - g0, ns,
                                              blk.stride(0) if g1 - g0 > 1 else ns, Qt.data_ptr() if rank else None, rank,
                                              Qt.stride(0) if rank > 1 else ns, cf.data_ptr(), cf.stride(0),
                                              mean.contiguous().data_ptr(), istd.data_ptr(), part.data_ptr()), "nsr_colvar")
                engine.LAUNCHES += 2
                col += part
            y = torch.log(torch.sqrt(col / nt))                                   # :106
            ym = y.mean()
            pred = ym + (Qc.T @ (Qc @ (y - ym)) if rc else 0.0)                   # :107-108
            new = torch.exp(pred)                                                 # :111
            if scale is not None:
                new = new * scale
            new = new / new.min()                                                 # :112
            prev, scale = scale, new
            it += 1
            if stepmax == 1:                                                      # no decision depends on t1: no sync
                best = scale
                break
            t1 = float(((new - prev) / prev).abs().max() if prev is not None else (new - 1.0).abs().max())      # :113
            if t1 < bestv:                                                        # :116-118
                bestv, best = t1, scale
        w = 1.0 / best                                                            # :122-123
        w = w / w.min()
        if not bool(torch.isfinite(w).all() & (w > 0).all()):
            raise AssertionError('non-finite variance normalisation multipliers')
        return w.cpu().numpy() if to_host else w
