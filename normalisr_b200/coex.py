"""``normalisr.coex.coex`` on the GPU (reference src/normalisr/coex.py:4-48)."""


def coex(dt, dc, **ka):
    """Co-expression for all gene pairs: ``(P, dot, var)``.

    dt: (n_gene, n_cell) normalised expression; dc: (n_cov, n_cell) covariates.
    P and dot are (n_gene, n_gene) with a zero diagonal, var is (n_gene,);
    Pearson R = dot / sqrt(var_i var_j)  (coex.py:30-34).
    Keyword arguments as the reference (``nth``, ``bsx``, ``bsy``, ``dimreduce`` ...; the
    batch-size / thread ones are accepted and ignored)."""
    from .association import association_tests
    ka.pop('bs', None)      # documented at coex.py:38; the reference itself rejects it downstream
    ans = association_tests(dt, None, dc, **ka)
    return (ans[0], ans[1], ans[4])
