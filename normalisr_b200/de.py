"""``normalisr.de.de`` on the GPU (reference src/normalisr/de.py:4-132)."""
import numpy as np
import torch


def _rows_that_vary(d):
    """len(np.unique(x)) > 1 for every row (de.py:92-93) without sorting and, for all but constant rows, without
    reading the row to its end: a row varies iff some entry differs from its first; column windows grow
    geometrically and only the rows still undecided are looked at again (a grouping indicator is decided in the
    first window, so the check costs microseconds instead of two passes over the whole host matrix)."""
    rows, n = d.shape
    keep = np.zeros(rows, dtype=bool)
    todo = np.arange(rows)
    c0, width = 0, 1024
    while todo.size and c0 < n:
        blk = d[:, c0:c0 + width] if todo.size == rows else d[todo, c0:c0 + width]
        first = d[:, :1] if todo.size == rows else d[todo, :1]
        v = (blk != first).any(axis=1)
        keep[todo[v]] = True
        todo = todo[~v]
        c0 += width
        width *= 4
    return keep


def de(dg, dt, dc, bs=0, **ka):
    """Differential expression of every gene against every grouping:
    ``(P, gamma, alpha|None, varg, vart)`` with the reference's shapes and fill values
    (groupings with a single value are skipped: P = 1, everything else 0; de.py:92-122).
    ``single`` = 0 (default), 1 (low MOI: every grouping tested on the cells that carry only it or
    nothing) or 4 ("other groupings as covariates")."""
    from .association import association_tests
    on_dev = isinstance(dg, torch.Tensor) and dg.is_cuda
    dg0 = dg
    if on_dev:
        keep = (dg0.max(dim=1).values != dg0.min(dim=1).values).cpu().numpy()
    else:
        if isinstance(dg0, torch.Tensor):
            dg0 = dg0.numpy()
        dg0 = np.asarray(dg0)
        keep = _rows_that_vary(dg0)
    if on_dev:
        dgk = dg0 if bool(keep.all()) else dg0[torch.from_numpy(keep).to(dg0.device)]
    else:
        dgk = dg0 if bool(keep.all()) else dg0[keep]
    if not on_dev and dgk.dtype != np.float64:
        dgk = dgk.astype(np.float64)
    P, gam, alpha, vg, vt = association_tests(dgk, dt, dc, bsx=bs, bsy=bs, return_dot=False, **ka)
    ng, nt, nc = dg0.shape[0], dt.shape[0], dc.shape[0]
    if on_dev and bool(keep.all()):
        # nothing to scatter: single=0 returns the genes' variance once, the reference's output has it per grouping
        return (P, gam, alpha, vg, vt if vt.dim() == 2 else vt[None, :].expand(ng, nt).contiguous())
    if on_dev:
        dev = dg0.device
        idx = torch.from_numpy(keep).to(dev)
        Pf = torch.ones((ng, nt), dtype=torch.float64, device=dev)
        gf = torch.zeros((ng, nt), dtype=torch.float64, device=dev)
        vgf = torch.zeros(ng, dtype=torch.float64, device=dev)
        vtf = torch.zeros((ng, nt), dtype=torch.float64, device=dev)
        Pf[idx], gf[idx], vgf[idx], vtf[idx] = P, gam, vg, vt
        af = None
        if alpha is not None:
            af = torch.zeros((ng, nt, nc), dtype=torch.float64, device=dev)
            af[idx] = alpha
        return (Pf, gf, af, vgf, vtf)
    odt = dt.dtype if isinstance(dt, np.ndarray) else np.float64
    if bool(keep.all()) and odt == np.float64:
        # nothing to scatter: the arrays that came back are the result (single=0 returns the genes' variance once,
        # the reference's output has it per grouping)
        return (P, gam, alpha, vg, vt if vt.ndim == 2 else np.repeat(vt[None, :], ng, axis=0))
    Pf = np.ones((ng, nt), dtype=odt)
    Pf[keep] = P
    gf = np.zeros((ng, nt), dtype=odt)
    gf[keep] = gam
    af = None
    if alpha is not None:
        af = np.zeros((ng, nt, nc), dtype=odt)
        af[keep] = alpha
    vgf = np.zeros(ng, dtype=odt)
    vgf[keep] = vg
    vtf = np.zeros((ng, nt), dtype=odt)
    vtf[keep] = vt                 # (nt,) broadcasts over groupings for single=0 (de.py:120-121)
    return (Pf, gf, af, vgf, vtf)
