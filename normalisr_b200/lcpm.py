"""``normalisr.lcpm.lcpm`` on the GPU (reference src/normalisr/lcpm.py:21-208):
Bayesian logCPM of a read-count matrix, the posterior mean digamma(1 + reads) - digamma(total + 2)
normalised per cell to log counts per million, plus the three cellular covariates (log total
reads, number of zero-count genes, its square), and optionally the reference's posterior
resampling (``varscale != 0``).

Streaming passes over the counts: ``nsr_lcpm_scan`` (min / max / total in one read: the negativity
check, the table length and the total), ``nsr_lcpm_colstats`` (per-cell sums; a pure table gather
without resampling) and ``nsr_lcpm_apply`` (gather + shift).  The digamma / trigamma look-up tables
over the count values (the reference builds the same tables) come from ``torch.special`` on the
device.  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.

Resampling draws one standard normal deviate per entry.  For numpy input they come from numpy's
global generator exactly as in the reference (``numpy.random.seed(seed)`` when a seed is given, then
one ``numpy.random.randn(n_gene, n_cell)``), so the result is the reference's, deviate for deviate;
for CUDA input they are generated on the device by Philox4x32-10 keyed by ``seed`` (a fresh random
key when ``seed`` is None) with the entry's index as counter - same distribution, another stream;
``noise=`` (a matrix of deviates) overrides both."""
import logging

import numpy as np
import torch

from . import _lib, engine, hoststage

_ROW_CHUNK_BYTES = 1 << 30
_HOLD_BYTES = 1 << 35          # host counts up to this size are copied to the device once for both passes


def _is_dev(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


_TABLE_LEN = 4096
_TABLES = {}


def _tables(dev):
    """(digamma(1 + c), exp(digamma(1 + c))) for c < 4096 on ``dev``."""
    key = torch.device(dev).index
    if key not in _TABLES:
        lut = torch.special.digamma(torch.arange(1, _TABLE_LEN + 1, dtype=torch.float64, device=dev)).contiguous()
        _TABLES[key] = (lut, torch.exp(lut).contiguous())
    return _TABLES[key]


def _trigamma(x):
    """polygamma(1, x) for x >= 1 to float64 rounding: 15 recurrence steps, then the asymptotic series up to
    B14 at x + 15 (torch.special.polygamma stops at B6 after 6 steps: 5e-10 relative, visible in the
    resampled values at the golden tolerance)."""
    s = torch.zeros_like(x)
    for i in range(15):
        s = s + 1.0 / ((x + i) * (x + i))
    iy = 1.0 / (x + 15.0)
    i2 = iy * iy
    ser = i2 * (1 / 6 + i2 * (-1 / 30 + i2 * (1 / 42 + i2 * (-1 / 30 + i2 * (5 / 66 + i2 * (-691 / 2730 + i2 * (7 / 6)))))))
    return s + iy * (1.0 + 0.5 * iy + ser)


def lcpm(reads, normalize=True, nth=0, ntot=None, varscale=0, seed=None, lowmem=True, nocov=False, device=None,
         noise=None):
    """Computes Bayesian log CPM from raw read counts: ``(lcpm, mean|None, var|None, cov|None)``
    with the reference's shapes (lcpm.py:29-66).  ``nth`` is accepted and ignored."""
    if reads.ndim != 2:
        raise ValueError('reads must have 2 dimensions.')
    if varscale < 0:
        raise ValueError('varscale must be non-negative.')
    if not normalize or ntot is not None or varscale != 0:
        logging.warning("Modifying keyword arguments other than nth or seed is neither recommended nor supported "
                        "for function 'lcpm'. Do so at your own risk.")
    to_host = not _is_dev(reads)
    ctx = engine.context(device if device is not None else (reads.device if not to_host else None))
    dev = ctx.device
    if to_host:
        if seed is not None:
            np.random.seed(seed)                            # lcpm.py:86-87
        if hasattr(reads, 'toarray'):                       # scipy sparse
            reads = reads.toarray()
        r = np.ascontiguousarray(reads)
        if not np.issubdtype(r.dtype, np.integer):
            r = r.astype(np.int64)                          # lcpm.py:129, 147
        if r.dtype not in (np.int32, np.int64):
            r = r.astype(np.int64 if r.dtype.itemsize > 4 or r.dtype.kind == 'u' and r.dtype.itemsize == 4 else np.int32)
        src = torch.from_numpy(r)
    else:
        src = reads
        if src.dtype not in (torch.int32, torch.int64):
            src = src.to(torch.int64)
        if src.stride(1) != 1:
            src = src.contiguous()
    nt, nc = src.shape
    itemsize = src.element_size()
    resample = varscale != 0
    with torch.cuda.device(dev):
        step = max(1, min(nt, min(_ROW_CHUNK_BYTES, hoststage.STAGE_BYTES) // max(1, 8 * nc))) if to_host else nt
        blocks = [(g0, min(nt, g0 + step)) for g0 in range(0, nt, step)]
        single = len(blocks) == 1
        held = {}
        # both passes read the counts: a host matrix that fits comfortably stays on the device in between
        keep_on_device = to_host and nt * nc * itemsize <= _HOLD_BYTES
        n_out = 1 + (0 if lowmem else 2)
        stage = hoststage.RowBlocks(ctx, src, step, out_row_bytes=8 * nc * n_out, tag="lcpm") if to_host else None

        def block(i):
            g0, g1 = blocks[i]
            if i in held:
                return held[i]
            blk = stage.fetch(g0, g1) if to_host else src[g0:g1]
            if single or keep_on_device:
                held[i] = blk
            return blk

        i64 = torch.iinfo(torch.int64)
        minmax = torch.tensor([i64.max, i64.min], dtype=torch.int64, device=dev)
        lut_var = lut_sd = None
        noise_d = None
        key = 0
        if resample:
            # the standard-deviation table must cover every count and needs the total: one extra read
            # (pass 0: min / max / total)
            scan = torch.empty(3, dtype=torch.int64, device=dev)
            acc = None
            for i, (g0, g1) in enumerate(blocks):
                blk = block(i)
                _lib.check(ctx.lib.nsr_lcpm_scan(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc,
                                                 blk.stride(0) if g1 - g0 > 1 else nc, scan.data_ptr()), "nsr_lcpm_scan")
                engine.LAUNCHES += 1
                cur = scan.cpu().numpy().copy() if nt and nc else np.array([0, 0, 0])
                acc = cur if acc is None else np.array([min(acc[0], cur[0]), max(acc[1], cur[1]), acc[2] + cur[2]])
            if acc is None:
                acc = np.array([0, 0, 0])
            if acc[0] < 0:
                raise ValueError('Negative value in d detected.')                # lcpm.py:88-89
            t0 = (int(acc[2]) if ntot is None else int(ntot)) + 2
            assert t0 > 2
            counts1 = torch.arange(1, int(acc[1]) + 2, dtype=torch.float64, device=dev)
            lut = torch.special.digamma(counts1).contiguous()                                    # :96-103 (+ const)
            lut_exp = None
            t0_t = torch.tensor(float(t0), dtype=torch.float64, device=dev)
            lut_var = (_trigamma(counts1) - _trigamma(t0_t)).contiguous()                        # :104-109
            lut_sd = torch.sqrt(lut_var * float(varscale)).contiguous()                          # :148-150
            if noise is None and to_host:
                noise = np.random.randn(nt, nc)             # the reference's own stream (:150)
            if noise is not None:
                noise_d = noise if isinstance(noise, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64))
                if tuple(noise_d.shape) != (nt, nc):
                    raise ValueError('noise must have the shape of reads.')
            else:
                key = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            # digamma(1 + c) and its exponential for the counts a table serves (larger ones are evaluated in
            # the kernels): independent of the data, built once per device
            lut, lut_exp = _tables(dev)

        def noise_args(g0, g1):
            if not resample:
                return None, None, 0, 0
            if noise_d is None:
                return lut_sd.data_ptr(), None, 0, key
            nb = noise_d[g0:g1].to(dev, torch.float64, non_blocking=True)
            if nb.stride(1) != 1:
                nb = nb.contiguous()
            keep.append(nb)
            return lut_sd.data_ptr(), nb.data_ptr(), nb.stride(0) if g1 - g0 > 1 else nc, 0

        keep = []
        # pass 1: per-cell statistics (+ smallest negative / largest count)
        col = torch.zeros((3, nc), dtype=torch.float64, device=dev)
        for i, (g0, g1) in enumerate(blocks):
            blk = block(i)
            part = torch.empty((3, nc), dtype=torch.float64, device=dev)
            sd_p, nz_p, ldn, sd_key = noise_args(g0, g1)
            _lib.check(ctx.lib.nsr_lcpm_colstats(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc,
                                                 blk.stride(0) if g1 - g0 > 1 else nc, lut.data_ptr(),
                                                 lut_exp.data_ptr() if lut_exp is not None else None,
                                                 lut.numel(), sd_p, nz_p, ldn, sd_key, g0, part.data_ptr(),
                                                 minmax.data_ptr()),
                       "nsr_lcpm_colstats")
            engine.LAUNCHES += 2
            col += part
        # digamma(total + 2), the constant of the reference's table: on the device, no host round trip
        total_t = col[1].sum() if ntot is None else torch.tensor(float(ntot), dtype=torch.float64, device=dev)
        psi_t0 = torch.special.digamma(total_t + 2.0)
        shift = psi_t0.expand(nc)                            # value = lut[c] - shift
        if normalize:                                        # lcpm.py:155-157: the constant cancels
            shift = torch.log(col[0]) - float(np.log(1e6))
        shift = shift.contiguous()
        dcov = None
        if not nocov:                                        # lcpm.py:193-199
            t1 = torch.log(col[1])
            dcov = torch.stack([t1, nt - col[2], t1 ** 2])
        # one small device->host read for the checks the reference makes up front
        chk = torch.stack([minmax[0].to(torch.float64), total_t.to(torch.float64), (col[1] == 0).any().to(torch.float64),
                           torch.isfinite(shift).all().to(torch.float64)]).cpu().numpy()
        if chk[0] < 0:
            raise ValueError('Negative value in d detected.')                    # lcpm.py:88-89
        assert chk[1] + 2 > 2                                                    # :91
        if not nocov and chk[2]:
            raise ValueError('Found cell with no read at all. Please remove.')   # :195-196
        if not chk[3]:
            raise AssertionError('non-finite logCPM')                            # :201
        # pass 2: gather (+ resampling) + per-cell shift
        host_out = to_host
        out = torch.empty((nt, nc), dtype=torch.float64, device=dev) if not host_out else torch.empty((nt, nc), dtype=torch.float64)
        dmean = dvar = None
        if not lowmem:                                       # lcpm.py:160-190
            dmean = torch.empty_like(out)
            dvar = torch.zeros_like(out)
        for i, (g0, g1) in enumerate(blocks):
            blk = block(i)
            res = out[g0:g1] if not host_out else torch.empty((g1 - g0, nc), dtype=torch.float64, device=dev)
            sd_p, nz_p, ldn, sd_key = noise_args(g0, g1)
            ld_b = blk.stride(0) if g1 - g0 > 1 else nc
            _lib.check(ctx.lib.nsr_lcpm_apply(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc, ld_b,
                                              lut.data_ptr(), lut.numel(), sd_p, nz_p, ldn, sd_key, g0,
                                              shift.data_ptr() if shift is not None else None, res.data_ptr(), nc),
                       "nsr_lcpm_apply")
            engine.LAUNCHES += 1
            if host_out:
                stage.store(out[g0:g1], res)
            if not lowmem:
                if resample:
                    m = torch.empty((g1 - g0, nc), dtype=torch.float64, device=dev)
                    _lib.check(ctx.lib.nsr_lcpm_apply(ctx.handle, engine._stream(), blk.data_ptr(), itemsize, g1 - g0, nc,
                                                      ld_b, lut.data_ptr(), lut.numel(), None, None, 0, 0, g0,
                                                      shift.data_ptr() if shift is not None else None, m.data_ptr(), nc),
                               "nsr_lcpm_apply")
                    engine.LAUNCHES += 1
                    v = lut_var[blk.long().clamp_(0, lut_var.numel() - 1)] * float(varscale)
                else:
                    m, v = res, None
                if host_out:
                    stage.store(dmean[g0:g1], m)
                    if v is not None:
                        stage.store(dvar[g0:g1], v)
                else:
                    dmean[g0:g1] = m
                    if v is not None:
                        dvar[g0:g1] = v
        if to_host:
            stage.close()
            torch.cuda.current_stream().synchronize()
            return (out.numpy(), None if dmean is None else dmean.numpy(), None if dvar is None else dvar.numpy(),
                    None if dcov is None else dcov.cpu().numpy())
        return (out, dmean, dvar, dcov)
