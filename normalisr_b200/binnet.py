"""Mirror of ``normalisr.binnet`` (reference ``src/normalisr/binnet.py``): P-value co-expression
network -> per-row Benjamini-Hochberg Q-values -> thresholded binary network.  This is the
immediate consumer of ``coex`` (reference CLI ``run.binnet``, run.py:313-321): with
P left on the device by ``normalisr_b200.normalisr.coex`` it returns a 1-byte/entry network
instead of shipping 8-byte P-values to the host.

    binnet(net, qcut) -> bool (n_gene, n_gene)         binnet.py:134-170
    bh(pv, weight=None) -> Q-values                    binnet.py:77-131

numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.  The per-row procedure runs in
``nsr_binnet`` (csrc/binnet.cu) and reproduces the reference's booleans bit for bit.
"""
import numpy as np
import torch

from . import _lib, engine, hoststage

_ROW_CHUNK_BYTES = 1 << 29


def _is_dev(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def binnet_rows(ctx, P, qcut, diag0=0, out=None, stats=None):
    """Rows of a P-value matrix (CUDA float64, unit column stride) -> uint8 network rows.
    diag0: column index of row 0's diagonal entry (rows of a row block of a larger matrix).
    stats: CUDA uint64[2] accumulating (edges, rows with entries outside [0, 1])."""
    assert P.is_cuda and P.dtype == torch.float64 and P.dim() == 2 and P.stride(1) == 1
    rows, cols = P.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.uint8, device=P.device)
    assert out.dtype == torch.uint8 and out.shape == P.shape and out.stride(1) == 1
    if stats is None:
        stats = torch.zeros(2, dtype=torch.int64, device=P.device)
    ld = P.stride(0) if rows > 1 else cols
    ldo = out.stride(0) if rows > 1 else cols
    _lib.check(ctx.lib.nsr_binnet(ctx.handle, engine._stream(), P.data_ptr(), rows, cols, ld, int(diag0), float(qcut),
                                  out.data_ptr(), ldo, stats.data_ptr()), "nsr_binnet")
    engine.LAUNCHES += 1
    return out, stats


def binnet(net, qcut, device=None):
    """Binarizes a P-value co-expression network to a thresholded Q-value network
    (binnet.py:134-170): Q-values per row (diagonal excluded), ``Q <= qcut``; the diagonal is False.
    Raises like the reference: ValueError (shape, qcut), AssertionError (values outside [0, 1] or
    not finite), RuntimeError("Empty binary network.")."""
    if net.ndim != 2:
        raise AssertionError('net must be 2-dimensional.')
    nt = net.shape[0]
    if net.shape[1] != nt or nt <= 1:
        raise ValueError('Wrong shape of net or namet.')
    if qcut <= 0 or qcut >= 1:
        raise ValueError('Q-value cutoff must be between 0 and 1.')
    to_host = not _is_dev(net)
    ctx = engine.context(device if device is not None else (net.device if not to_host else None))
    with torch.cuda.device(ctx.device):
        stats = torch.zeros(2, dtype=torch.int64, device=ctx.device)
        if not to_host:
            Pd = net.to(torch.float64)
            if Pd.stride(1) != 1:
                Pd = Pd.contiguous()
            out, _ = binnet_rows(ctx, Pd, qcut, 0, stats=stats)
        else:
            # rows are independent: stage the host matrix in row chunks
            src = net if isinstance(net, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(net))
            if src.dtype != torch.float64:
                src = src.to(torch.float64)
            out = torch.empty((nt, nt), dtype=torch.uint8, device=ctx.device)
            step = max(1, min(nt, min(_ROW_CHUNK_BYTES, hoststage.STAGE_BYTES) // (8 * nt)))
            blocks = hoststage.RowBlocks(ctx, src, step, tag="binnet")       # pageable P through page-locked slots
            for r0 in range(0, nt, step):
                r1 = min(nt, r0 + step)
                binnet_rows(ctx, blocks.fetch(r0, r1), qcut, r0, out=out[r0:r1], stats=stats)
        edges, invalid = (int(x) for x in stats.cpu())
        if invalid:
            raise AssertionError('net must be finite with values in [0, 1].')
        if edges == 0:
            raise RuntimeError("Empty binary network.")
        out = out.view(torch.bool)
        if to_host:
            from .association import _outs
            return _outs((out,))[0]
        return out


def bh(pv, weight=None, device=None):
    """Benjamini-Hochberg Q-values of a vector (binnet.py:77-131): unique P-values, cumulative
    (weighted) counts normalised to 1, q = p / w clipped to [0, 1], running minimum from the top.
    A helper next to ``binnet`` (which does not call it: the row kernel needs no sort); runs on
    the device with torch primitives."""
    to_host = not _is_dev(pv)
    ctx = engine.context(device if device is not None else (pv.device if not to_host else None))
    p = (pv if isinstance(pv, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pv))).to(ctx.device)
    assert p.dim() == 1 and p.numel() > 0
    assert bool(torch.isfinite(p).all()) and float(p.min()) >= 0 and float(p.max()) <= 1
    if weight is None:
        wt = torch.ones_like(p, dtype=torch.float64)
    else:
        wt = (weight if isinstance(weight, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(weight))).to(
            ctx.device, torch.float64)
        assert wt.shape == p.shape
        assert bool(torch.isfinite(wt).all()) and float(wt.min()) >= 0 and float(wt.max()) > 0
    vals, inv = torch.unique(p, return_inverse=True)
    w = torch.zeros(vals.numel(), dtype=torch.float64, device=ctx.device).index_add_(0, inv, wt)
    w = torch.cumsum(w, 0)
    w = w / w[-1]
    q = vals.to(torch.float64) / w
    q = torch.where(torch.isfinite(q), q, torch.ones_like(q)).clamp_(0, 1)
    q = torch.flip(torch.cummin(torch.flip(q, [0]), 0).values, [0])
    ans = q[inv].to(p.dtype)
    return ans.cpu().numpy() if to_host else ans


def nodiag(d, split=False):
    """Removes the diagonal of a 2-D numpy array (binnet.py:4-33): the rows without their diagonal
    entry, as a list (split) or concatenated."""
    d = np.asarray(d)
    assert d.ndim == 2
    k = min(d.shape)
    rows = [np.concatenate([d[i, :i], d[i, i + 1:]]) if i < k else d[i] for i in range(d.shape[0])]
    return rows if split else np.concatenate(rows)


def rediag(d, fill=0, shape=None):
    """Inverse of ``nodiag(split=False)`` (binnet.py:36-74)."""
    d = np.asarray(d)
    if shape is None:
        t1 = int(np.sqrt(d.size)) + 1
        assert t1 * (t1 - 1) == d.size
        shape = (t1, t1)
    else:
        assert len(shape) == 2 and shape[0] * shape[1] - min(shape) == d.size
    m = np.full(shape, fill, dtype=d.dtype)
    k = min(shape)
    mask = np.ones(shape, dtype=bool)
    mask[np.arange(k), np.arange(k)] = False
    m[mask] = d
    return m
