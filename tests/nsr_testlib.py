"""numpy restatements of the device-side data transforms (test helpers only)."""
import numpy as np


SIGN_MASK = 0x9e3779b9


def cell_flip(k):
    """residual.cu reg_flip(): cell m = k % 128 of every block is negated when bit 2 (m // 8) + m % 2 of
    kSignMask is set (the same pattern for every block)."""
    m = np.asarray(k, dtype=np.int64) % 128
    i = 2 * (m >> 3) + (m & 1)
    return ((SIGN_MASK >> i) & 1).astype(bool)


def hadamard128(z):
    """Sign-randomised orthonormal 128-point Walsh-Hadamard transform along the last axis
    (natural/Hadamard ordering), cells zero-padded to a multiple of 128."""
    rows, n = z.shape
    n_pad = (n + 127) // 128 * 128
    zp = np.zeros((rows, n_pad))
    zp[:, :n] = z
    sign = np.where(cell_flip(np.arange(n_pad)), -1.0, 1.0)
    zp = zp * sign[None, :]
    x = zp.reshape(rows, n_pad // 128, 128)
    # device element index e = 32*j + lane: butterflies over all 7 bits -> plain WHT on index e
    h = 1
    while h < 128:
        x = x.reshape(rows, -1, 128 // (2 * h), 2, h)
        a = x[:, :, :, 0, :] + x[:, :, :, 1, :]
        b = x[:, :, :, 0, :] - x[:, :, :, 1, :]
        x = np.stack([a, b], axis=3).reshape(rows, -1, 128)
        h *= 2
    return (x / np.sqrt(128.0)).reshape(rows, n_pad)


def residual(x, dc_basis):
    """x - (x Qt^T) Qt."""
    if dc_basis is None or dc_basis.shape[0] == 0:
        return x.copy()
    return x - (x @ dc_basis.T) @ dc_basis


def digits_to_int(slices):
    """(S, rows, n_pad) int8 -> int64 value per element."""
    v = np.zeros(slices.shape[1:], dtype=np.int64)
    for s in range(slices.shape[0]):
        v = v * 256 + slices[s].astype(np.int64)
    return v


def kept_products_sum(slices_a, slices_b, wmax):
    """Sum over cells of the kept digit-pair products (a+b <= wmax, 1-based) with their
    base-256 weights: exact int64 per pair, combined in long double."""
    S = slices_a.shape[0]
    tot = None
    for a in range(S):
        for b in range(S):
            if a + b + 2 <= wmax:
                w = np.longdouble(256) ** (2 * S - (a + b + 2))
                p = (slices_a[a].astype(np.int64) @ slices_b[b].astype(np.int64).T).astype(np.longdouble)
                tot = p * w if tot is None else tot + p * w
    return tot
