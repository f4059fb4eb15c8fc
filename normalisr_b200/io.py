"""Text I/O of the command-line layer (reference ``src/normalisr/run.py:10-35``): same function names and
argument meaning as ``run.file_read_tsv`` / ``file_write_tsv`` / ``file_read_coo``, with the parsing done by
the native library's thread pool straight into page-locked buffers (``nsr_tsv_*`` / ``nsr_mtx_*``), so a
matrix goes to the device with one asynchronous copy.

``device=None`` returns numpy arrays (scipy COO for ``file_read_coo``) like the reference;
``device='cuda'`` (or an index) returns CUDA tensors - for count matrices a dense int32 tensor, the
input ``lcpm`` takes.  These functions need no GPU unless a device is asked for."""
import ctypes
import logging
import os

import numpy as np

from . import _lib

fmt_float = '%.8G'          # run.py:6
fmt_int = '%i'


def _pinned(shape, dtype):
    """Page-locked host tensor when CUDA is usable (asynchronous copies), ordinary memory otherwise."""
    import torch
    try:
        if torch.cuda.is_available():
            return torch.empty(shape, dtype=dtype, pin_memory=True)
    except Exception:
        pass
    return torch.empty(shape, dtype=dtype)


def file_read_tsv(f, delimiter='\t', device=None, nth=0, **ka):
    """Read a table of numbers (run.py:20-27, ``numpy.loadtxt(f, delimiter=delimiter)``): float64
    (rows, cols), a single row stays 2-D."""
    if ka:
        raise TypeError('file_read_tsv: unsupported keyword arguments {} (plain numeric tables only)'.format(sorted(ka)))
    import torch
    lib = _lib.load()
    logging.debug('Start reading file ' + f)
    rows, cols = ctypes.c_int64(), ctypes.c_int64()
    d = delimiter.encode() if isinstance(delimiter, str) else delimiter
    _lib.check(lib.nsr_tsv_shape(os.fsencode(f), d, ctypes.byref(rows), ctypes.byref(cols)), "nsr_tsv_shape")
    buf = _pinned((rows.value, cols.value), torch.float64)
    if rows.value and cols.value:
        _lib.check(lib.nsr_tsv_read(os.fsencode(f), d, buf.data_ptr(), rows.value, cols.value, cols.value, int(nth)),
                   "nsr_tsv_read")
    logging.debug('Finish reading file ' + f)
    if device is not None:
        return buf.to(torch.device('cuda', device) if isinstance(device, int) else device, non_blocking=True)
    return buf.numpy()


def file_write_tsv(f, d, delimiter='\t', fmt=fmt_float, nth=0, **ka):
    """Write a table (run.py:30-35, ``numpy.savetxt(f, d, delimiter=delimiter, fmt=fmt)``): the same bytes.
    ``fmt`` must be of the form '%.<k>G' (the reference's '%.8G') or '%i'."""
    import torch
    if ka:
        raise TypeError('file_write_tsv: unsupported keyword arguments {}'.format(sorted(ka)))
    if isinstance(d, torch.Tensor):
        d = d.detach().cpu().numpy()
    d = np.asarray(d)
    if d.ndim == 1:
        d = d.reshape(-1, 1)                                 # numpy.savetxt writes a 1-D array as a column
    if d.ndim != 2:
        raise ValueError('Expected 1D or 2D array, got %dD array instead' % d.ndim)
    if fmt == fmt_int:
        return np.savetxt(f, d, delimiter=delimiter, fmt=fmt)
    if not (fmt.startswith('%.') and fmt.endswith('G') and fmt[2:-1].isdigit()):
        raise ValueError("file_write_tsv: fmt must be '%.<k>G' or '%i'")
    x = np.ascontiguousarray(d, dtype=np.float64)
    lib = _lib.load()
    logging.debug('Start writing file ' + f)
    dl = delimiter.encode() if isinstance(delimiter, str) else delimiter
    _lib.check(lib.nsr_tsv_write(os.fsencode(f), dl, x.ctypes.data, x.shape[0], x.shape[1], x.shape[1], int(fmt[2:-1]),
                                 int(nth)), "nsr_tsv_write")
    logging.debug('Finish writing file ' + f)


def file_read_coo(f, device=None, nth=0, **ka):
    """Read a MatrixMarket coordinate file (run.py:10-17, ``scipy.io.mmread``).  ``device=None``: a
    ``scipy.sparse.coo_matrix`` like the reference (integer files give an integer matrix);
    with a device: the DENSE matrix as a CUDA tensor (int32 for integer files, float64 otherwise) -
    the read-count matrix ``lcpm`` takes."""
    if ka:
        raise TypeError('file_read_coo: unsupported keyword arguments {}'.format(sorted(ka)))
    import torch
    lib = _lib.load()
    logging.debug('Start reading file ' + f)
    rows, cols, nnz, is_int = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    _lib.check(lib.nsr_mtx_shape(os.fsencode(f), ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(nnz),
                                 ctypes.byref(is_int)), "nsr_mtx_shape")
    n = nnz.value
    r = _pinned((n,), torch.int32)
    c = _pinned((n,), torch.int32)
    v = _pinned((n,), torch.float64)
    sym = ctypes.c_int(0)
    if n:
        _lib.check(lib.nsr_mtx_read(os.fsencode(f), n, r.data_ptr(), c.data_ptr(), v.data_ptr(), ctypes.byref(sym), int(nth)),
                   "nsr_mtx_read")
    logging.debug('Finish reading file ' + f)
    if device is None:
        import scipy.sparse
        rr, cc, vv = r.numpy(), c.numpy(), v.numpy()
        if sym.value:                                        # mirror the off-diagonal entries, as scipy does
            off = rr != cc
            rr, cc, vv = (np.concatenate([rr, cc[off]]), np.concatenate([cc, rr[off]]),
                          np.concatenate([vv, vv[off] * (-1.0 if sym.value == 2 else 1.0)]))
        if is_int.value:
            vv = vv.astype(np.int64)
        return scipy.sparse.coo_matrix((vv, (rr, cc)), shape=(rows.value, cols.value))
    dev = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
    rd, cd, vd = r.to(dev, non_blocking=True).long(), c.to(dev, non_blocking=True).long(), v.to(dev, non_blocking=True)
    out = torch.zeros((rows.value, cols.value), dtype=torch.float64, device=dev)
    out.index_put_((rd, cd), vd, accumulate=True)            # duplicate entries add up, as in scipy's conversion
    if sym.value:
        off = rd != cd
        out.index_put_((cd[off], rd[off]), vd[off] * (-1.0 if sym.value == 2 else 1.0), accumulate=True)
    return out.to(torch.int32) if is_int.value else out
