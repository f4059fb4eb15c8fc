"""``normalisr.de.de`` on the GPU (reference src/normalisr/de.py:4-132)."""
import numpy as np
import torch


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
        dg0 = np.asarray(dg0)
        keep = np.array([len(np.unique(x)) > 1 for x in dg0], dtype=bool)      # de.py:92-93
    dgk = dg0[torch.from_numpy(keep).to(dg0.device)] if on_dev else dg0[keep]
    if not on_dev and dgk.dtype != np.float64:
        dgk = dgk.astype(np.float64)
    P, gam, alpha, vg, vt = association_tests(dgk, dt, dc, bsx=bs, bsy=bs, return_dot=False, **ka)
    ng, nt, nc = dg0.shape[0], dt.shape[0], dc.shape[0]
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
