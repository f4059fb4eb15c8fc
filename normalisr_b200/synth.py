"""Synthetic inputs of the shapes BASELINE.json names (SURVEY.md 8(d) recipe).

`host_*` functions use numpy (tests, CPU baseline); `device_*` use torch on the GPU so the
100k x 20k matrix never has to exist on the host.  The expression matrix is "normvar-shaped":
log-normalised NB counts with a few planted co-expression modules, so P spans 1 .. <1e-300.
"""
import numpy as np


def host_counts(rng, n_gene, n_cell, n_module=0, module_size=30):
    """NB(size=2) counts with gene means Gamma(0.5, 2)+0.05 and LogNormal(0, 0.4) depths;
    optional planted modules share a per-cell log factor (so they are co-expressed)."""
    mu = rng.gamma(0.5, 2.0, size=n_gene) + 0.05
    depth = rng.lognormal(0.0, 0.4, size=n_cell)
    m = mu[:, None] * depth[None, :]
    for k in range(n_module):
        f = np.exp(rng.normal(0, rng.uniform(0.3, 1.0), size=n_cell))
        idx = rng.choice(n_gene, size=min(module_size, n_gene), replace=False)
        m[idx] = rng.uniform(3, 15, size=len(idx))[:, None] * depth[None, :] * f[None, :]
    return rng.negative_binomial(2, 2.0 / (2.0 + m)), depth


def host_problem(seed, n_gene, n_cell, n_batch=6, n_module=4, n_group=0, group_p=0.02):
    """Return dict(dt, dc[, dg]): float64, rows = variables, columns = cells.  dc = one-hot
    batches minus one + log depth + detection rate + a constant row (like normcov's output)."""
    rng = np.random.default_rng(seed)
    reads, depth = host_counts(rng, n_gene, n_cell, n_module)
    tot = reads.sum(axis=0) + 1.0
    dt = np.log((reads + 0.5) / tot[None, :] * 1e4 + 1.0)
    b = rng.integers(0, n_batch, size=n_cell)
    cov = [(b == i).astype(float) for i in range(1, n_batch)]
    ld = np.log(tot)
    det = (reads > 0).mean(axis=0)
    cov += [(ld - ld.mean()) / ld.std(), (det - det.mean()) / det.std()]
    dc = np.array(cov + [np.ones(n_cell)])
    out = {"dt": np.ascontiguousarray(dt), "dc": np.ascontiguousarray(dc)}
    if n_group:
        dg = (rng.random((n_group, n_cell)) < group_p).astype(np.float64)
        k = min(n_group, n_gene, 8)
        out["dt"][:k] += 0.5 * dg[:k]            # a few true effects
        out["dg"] = dg
    return out


def device_problem(seed, n_gene, n_cell, device, n_batch=6, n_module=8, module_size=200,
                   n_group=0, group_p=0.02, gene_chunk=2048):
    """Same recipe generated on the GPU with torch (plumbing only), chunked over genes."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    depth = torch.exp(0.4 * torch.randn(n_cell, generator=g, device=device, dtype=f64))
    mu = torch.distributions.Gamma(torch.tensor(0.5, device=device, dtype=f64),
                                   torch.tensor(0.5, device=device, dtype=f64))
    torch.manual_seed(seed)
    mu = mu.sample((n_gene,)) + 0.05
    factors = [torch.exp(torch.randn(n_cell, generator=g, device=device, dtype=f64) * (0.3 + 0.7 * k / max(1, n_module)))
               for k in range(n_module)]
    dt = torch.empty((n_gene, n_cell), dtype=f64, device=device)
    tot = torch.zeros(n_cell, dtype=f64, device=device)
    det = torch.zeros(n_cell, dtype=f64, device=device)
    for g0 in range(0, n_gene, gene_chunk):
        g1 = min(g0 + gene_chunk, n_gene)
        m = mu[g0:g1, None] * depth[None, :]
        for k, f in enumerate(factors):
            lo = (k * module_size) % max(1, n_gene)
            a, b_ = max(lo, g0), min(lo + module_size, g1)
            if a < b_:
                lvl = 3.0 + 12.0 * torch.rand(b_ - a, generator=g, device=device, dtype=f64)
                m[a - g0:b_ - g0] = lvl[:, None] * depth[None, :] * f[None, :]
        # NB(size=2) as a Gamma-Poisson mixture
        lam = torch.distributions.Gamma(torch.full_like(m, 2.0), 2.0 / m).sample()
        reads = torch.poisson(lam, generator=g)
        dt[g0:g1] = reads
        tot += reads.sum(dim=0)
        det += (reads > 0).sum(dim=0)
    tot += 1.0
    for g0 in range(0, n_gene, gene_chunk):
        g1 = min(g0 + gene_chunk, n_gene)
        dt[g0:g1] = torch.log((dt[g0:g1] + 0.5) / tot[None, :] * 1e4 + 1.0)
    b = torch.randint(0, n_batch, (n_cell,), generator=g, device=device)
    cov = [(b == i).to(f64) for i in range(1, n_batch)]
    ld = torch.log(tot)
    dr = det / n_gene
    cov += [(ld - ld.mean()) / ld.std(), (dr - dr.mean()) / dr.std(), torch.ones(n_cell, dtype=f64, device=device)]
    out = {"dt": dt, "dc": torch.stack(cov)}
    if n_group:
        dg = (torch.rand((n_group, n_cell), generator=g, device=device) < group_p).to(f64)
        k = min(n_group, n_gene, 8)
        dt[:k] += 0.5 * dg[:k]
        out["dg"] = dg
    return out
