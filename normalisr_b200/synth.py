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


def host_problem(seed, n_gene, n_cell, n_batch=6, n_module=4, n_group=0, group_p=0.02, module_size=30):
    """Return dict(dt, dc[, dg]): float64, rows = variables, columns = cells.  dc = one-hot
    batches minus one + 3 continuous summaries + a constant row: 9 covariates, like the
    reference chain's 5 batch indicators + lcpm's 3 covariates + normcov's constant row."""
    rng = np.random.default_rng(seed)
    reads, depth = host_counts(rng, n_gene, n_cell, n_module, module_size)
    tot = reads.sum(axis=0) + 1.0
    dt = np.log((reads + 0.5) / tot[None, :] * 1e4 + 1.0)
    b = rng.integers(0, n_batch, size=n_cell)
    cov = [(b == i).astype(float) for i in range(1, n_batch)]
    ld = np.log(tot)
    det = (reads > 0).mean(axis=0)
    dv = np.log1p(reads).var(axis=0)
    cov += [(ld - ld.mean()) / ld.std(), (det - det.mean()) / det.std(), (dv - dv.mean()) / dv.std()]
    dc = np.array(cov + [np.ones(n_cell)])
    out = {"dt": np.ascontiguousarray(dt), "dc": np.ascontiguousarray(dc)}
    if n_group:
        dg = (rng.random((n_group, n_cell)) < group_p).astype(np.float64)
        k = min(n_group, n_gene, 8)
        out["dt"][:k] += 0.5 * dg[:k]            # a few true effects
        out["dg"] = dg
    return out


def device_problem(seed, n_gene, n_cell, device, n_batch=6, n_module=8, module_size=200,
                   n_group=0, group_p=0.02, gene_chunk=2048, gene_seed=None):
    """Same recipe generated on the GPU with torch (plumbing only), chunked over genes.
    Everything per cell (depth, batches, covariates, groupings) depends on ``seed`` only, so
    ranks that pass different ``gene_seed`` values get different genes of the same cells."""
    import torch
    f64 = torch.float64
    gc = torch.Generator(device=device)
    gc.manual_seed(seed)
    gg = torch.Generator(device=device)
    gg.manual_seed(seed * 7919 + 13 if gene_seed is None else gene_seed)
    depth = torch.exp(0.4 * torch.randn(n_cell, generator=gc, device=device, dtype=f64))
    b = torch.randint(0, n_batch, (n_cell,), generator=gc, device=device)
    noise = torch.randn(n_cell, generator=gc, device=device, dtype=f64)
    cov = [(b == i).to(f64) for i in range(1, n_batch)]
    ld = torch.log(depth)
    noise2 = torch.randn(n_cell, generator=gc, device=device, dtype=f64)
    c2 = 0.6 * ld + 0.8 * noise
    c3 = 0.3 * ld - 0.5 * noise + 0.8 * noise2
    cov += [(ld - ld.mean()) / ld.std(), (c2 - c2.mean()) / c2.std(), (c3 - c3.mean()) / c3.std(),
            torch.ones(n_cell, dtype=f64, device=device)]
    dc = torch.stack(cov)
    # gene means ~ Gamma(0.5, scale 2) + 0.05 via the square of a normal (chi2_1 = Gamma(1/2, 2))
    mu = torch.randn(n_gene, generator=gg, device=device, dtype=f64) ** 2 + 0.05
    factors = [torch.exp(torch.randn(n_cell, generator=gg, device=device, dtype=f64) *
                         (0.3 + 0.7 * k / max(1, n_module))) for k in range(n_module)]
    dt = torch.empty((n_gene, n_cell), dtype=f64, device=device)
    scale = 1e4 / (depth * (mu.sum() + 1.0))
    for g0 in range(0, n_gene, gene_chunk):
        g1 = min(g0 + gene_chunk, n_gene)
        m = mu[g0:g1, None] * depth[None, :]
        for k, f in enumerate(factors):
            lo = (k * module_size) % max(1, n_gene)
            a, b_ = max(lo, g0), min(lo + module_size, g1)
            if a < b_:
                lvl = 3.0 + 12.0 * torch.rand(b_ - a, generator=gg, device=device, dtype=f64)
                m[a - g0:b_ - g0] = lvl[:, None] * depth[None, :] * f[None, :]
        # NB(size=2) = Poisson with Gamma(2, m/2) rate; Gamma(2) = sum of two exponentials
        u = torch.rand((2,) + m.shape, generator=gg, device=device, dtype=f64).clamp_min_(1e-300)
        lam = -(torch.log(u[0]) + torch.log(u[1])) * (m * 0.5)
        reads = torch.poisson(lam, generator=gg)
        dt[g0:g1] = torch.log((reads + 0.5) * scale[None, :] + 1.0)
        del m, u, lam, reads
    out = {"dt": dt, "dc": dc}
    if n_group:
        dg = (torch.rand((n_group, n_cell), generator=gc, device=device) < group_p).to(f64)
        k = min(n_group, n_gene, 8)
        dt[:k] += 0.5 * dg[:k]
        out["dg"] = dg
    return out
