"""``normalisr.lcpm.lcpm`` on the GPU (reference src/normalisr/lcpm.py:21-208):
Bayesian logCPM of a read-count matrix, the posterior mean digamma(1 + reads) - digamma(total + 2)
normalised per cell to log counts per million, plus the three cellular covariates (log total
reads, number of zero-count genes, its square).

Two streaming passes over the counts (``nsr_lcpm_colstats``, ``nsr_lcpm_apply``); the digamma
look-up table over the count values (the reference builds the same table) comes from
``torch.special.digamma`` on the device.  numpy in -> numpy out, CUDA tensors in -> CUDA tensors
out.  Only the supported configuration of the reference is accelerated (``varscale=0``: no
posterior resampling)."""
import logging

import numpy as np
import torch

from . import _lib, engine

_ROW_CHUNK_BYTES = 1 << 30


def _is_dev(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def lcpm(reads, normalize=True, nth=0, ntot=None, varscale=0, seed=None, lowmem=True, nocov=False, device=None):
    """Computes Bayesian log CPM from raw read counts: ``(lcpm, mean|None, var|None, cov|None)``
    with the reference's shapes (lcpm.py:29-66).  ``nth`` / ``seed`` are accepted and ignored."""
    if reads.ndim != 2:
        raise ValueError('reads must have 2 dimensions.')
    if varscale < 0:
        raise ValueError('varscale must be non-negative.')
    if varscale != 0:
        raise NotImplementedError('normalisr_b200 accelerates lcpm without posterior resampling (varscale=0).')
    if not normalize or ntot is not None:
        logging.warning("Modifying keyword arguments other than nth or seed is neither recommended nor supported "
                        "for function 'lcpm'. Do so at your own risk.")
    to_host = not _is_dev(reads)
    ctx = engine.context(device if device is not None else (reads.device if not to_host else None))
    dev = ctx.device
    if to_host:
        if hasattr(reads, 'toarray'):                       # scipy sparse
            reads = reads.toarray()
        r = np.ascontiguousarray(reads)
        if not np.issubdtype(r.dtype, np.integer):
            r = r.astype(np.int64)                          # lcpm.py:129, 147
        if r.size and r.min() < 0:
            raise ValueError('Negative value in d detected.')
        if r.dtype not in (np.int32, np.int64):
            r = r.astype(np.int64 if r.dtype.itemsize > 4 or (r.size and int(r.max()) > 2**31 - 1) else np.int32)
        src = torch.from_numpy(r)
    else:
        src = reads
        if src.dtype not in (torch.int32, torch.int64):
            src = src.to(torch.int64)
        if src.stride(1) != 1:
            src = src.contiguous()
        if bool((src < 0).any()):
            raise ValueError('Negative value in d detected.')
    nt, nc = src.shape
    itemsize = src.element_size()
    with torch.cuda.device(dev):
        step = max(1, _ROW_CHUNK_BYTES // max(1, itemsize * nc)) if to_host else nt
        blocks = [(g0, min(nt, g0 + step)) for g0 in range(0, nt, step)]
        max_count = int(src.max()) if src.numel() else 0
        # pass 0 (only when the total is not given): total reads.  Needed before the table exists.
        if ntot is None:
            total = int(src.sum(dtype=torch.int64)) if not to_host else int(r.sum(dtype=np.int64))
        else:
            total = int(ntot)
        t0 = total + 2
        assert t0 > 2
        lut = (torch.special.digamma(torch.arange(1, max_count + 2, dtype=torch.float64, device=dev))
               - torch.special.digamma(torch.tensor(float(t0), dtype=torch.float64, device=dev)))
        lut = lut.contiguous()
        # pass 1: per-cell statistics
        col = torch.zeros((3, nc), dtype=torch.float64, device=dev)
        held = []
        for g0, g1 in blocks:
            blk = src[g0:g1].to(dev, non_blocking=True) if to_host else src[g0:g1]
            part = torch.empty((3, nc), dtype=torch.float64, device=dev)
            _lib.check(ctx.lib.nsr_lcpm_colstats(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc,
                                                 blk.stride(0) if g1 - g0 > 1 else nc, lut.data_ptr(), lut.numel(),
                                                 part.data_ptr()), "nsr_lcpm_colstats")
            engine.LAUNCHES += 2
            col += part
            if len(blocks) == 1:
                held.append(blk)
        shift = None
        if normalize:                                        # lcpm.py:155-157
            shift = (torch.log(col[0]) - float(np.log(1e6))).contiguous()
        dcov = None
        if not nocov:                                        # lcpm.py:193-199
            if bool((col[1] == 0).any()):
                raise ValueError('Found cell with no read at all. Please remove.')
            t1 = torch.log(col[1])
            dcov = torch.stack([t1, nt - col[2], t1 ** 2])
        # pass 2: gather + per-cell shift
        out = torch.empty((nt, nc), dtype=torch.float64, device=dev) if not to_host else torch.empty((nt, nc), dtype=torch.float64)
        for g0, g1 in blocks:
            blk = held[0] if held else (src[g0:g1].to(dev, non_blocking=True) if to_host else src[g0:g1])
            res = out[g0:g1] if not to_host else torch.empty((g1 - g0, nc), dtype=torch.float64, device=dev)
            _lib.check(ctx.lib.nsr_lcpm_apply(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc,
                                              blk.stride(0) if g1 - g0 > 1 else nc, lut.data_ptr(), lut.numel(),
                                              shift.data_ptr() if shift is not None else None, res.data_ptr(), nc),
                       "nsr_lcpm_apply")
            engine.LAUNCHES += 1
            if to_host:
                out[g0:g1] = res.cpu()
        if not bool(torch.isfinite(shift).all() if shift is not None else True):
            raise AssertionError('non-finite logCPM')        # lcpm.py:201
        dmean = dvar = None
        if not lowmem:                                       # lcpm.py:178-190 with varscale = 0
            dmean = out.clone()
            dvar = torch.zeros_like(out)
        if to_host:
            torch.cuda.current_stream().synchronize()
            return (out.numpy(), None if dmean is None else dmean.numpy(), None if dvar is None else dvar.numpy(),
                    None if dcov is None else dcov.cpu().numpy())
        return (out, dmean, dvar, dcov)
