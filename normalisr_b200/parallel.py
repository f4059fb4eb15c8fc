"""Multi-GPU co-expression / DE: one process per GPU, torch.distributed (NCCL) for the single
exchange step.  Replaces the reference's thread pool over tiles (``parallel.autopooler``,
parallel.py:12-74, used at association.py:997).

The path shards naturally:
  * every rank residualises and quantises its own block of genes (covariates are tiny and
    replicated);
  * ONE all-gather of the int8 digit planes (+ per-row quantum and variance);
  * each rank then owns a strip of 128-row output tile rows, balanced by tile count, and
    computes the upper triangle of that strip with no further communication.
The full matrices are U + U^T over the ranks' strips (``gather_dense`` assembles them).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import engine
from ._lib import MODE_COEX_UPPER, MODE_DE, TILE
from .association import covariate_basis


def row_split(rows, world):
    """Contiguous, equally sized row blocks (last one short): block size."""
    return (rows + world - 1) // world


def strip_bounds(n_tile_rows, world):
    """Tile-row strips [a_k, b_k) with (nearly) equal numbers of upper-triangular tiles.
    Tile row i holds n_tile_rows - i tiles."""
    total = n_tile_rows * (n_tile_rows + 1) // 2
    bounds = [0]
    acc, k = 0, 1
    for i in range(n_tile_rows):
        acc += n_tile_rows - i
        while k < world and acc >= total * k / world:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world:
        bounds.append(n_tile_rows)
    bounds.append(n_tile_rows)
    return [(bounds[k], max(bounds[k], bounds[k + 1])) for k in range(world)]


def strip_tiles(n_tile_rows, a, b, strip=12):
    """Upper-triangular tiles with tile row in [a, b), column-strip ordered for L2 reuse."""
    out = []
    for js in range(a, n_tile_rows, strip):
        je = min(js + strip, n_tile_rows)
        for i in range(a, min(b, je)):
            for j in range(max(i, js), je):
                out.append((i, j))
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def gather_sliced(local, rows_total, group=None):
    """All-gather equally sized row blocks of digit planes into one Sliced of rows_total rows."""
    world = dist.get_world_size(group)
    blk = local.rows_alloc
    full = engine.Sliced(blk * world, local.n, local.n_slices, local.slices.device)
    for s in range(local.n_slices):
        dist.all_gather_into_tensor(full.slices[s], local.slices[s].contiguous(), group=group)
    dist.all_gather_into_tensor(full.quantum, local.quantum, group=group)
    dist.all_gather_into_tensor(full.var, local.var, group=group)
    full.energy_max.copy_(local.energy_max)
    dist.all_reduce(full.energy_max, op=dist.ReduceOp.MAX, group=group)   # maxima over all ranks' rows
    full.rows = rows_total            # rows beyond rows_total are padding of the last block
    return full


def residualize_block(ctx, x_block, Qt_dev, n_slices, blk):
    """Residualise this rank's rows into a block padded to ``blk`` rows (padding rows are zero
    planes with quantum 1, var 1)."""
    out = engine.Sliced(blk, x_block.shape[1], n_slices, ctx.device)
    rows = x_block.shape[0]
    if rows < blk:
        out.slices[:, rows:].zero_()
        out.quantum[rows:] = 1.0
        out.var[rows:] = 1.0
    if rows:
        engine.residualize(ctx, x_block, Qt_dev, n_slices, out=out, row_offset=0)
    return out


def coex_sharded(dt_block, dc, n_gene, group=None, precision="default", dimreduce=0, out=None):
    """Co-expression over all ranks of ``group``.

    dt_block: this rank's genes, rows [rank*blk, min((rank+1)*blk, n_gene)) of the expression
              matrix, blk = row_split(n_gene, world); CUDA float64 (rows_local, n_cell).
    dc:       full covariate matrix (numpy or tensor), identical on all ranks.
    Returns (P_strip, dot_strip, var, (row_begin, row_end)): the upper triangle of this rank's
    strip of rows (entries left of the diagonal tile are not written), var for all genes.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ctx = engine.context(dt_block.device)
    n_slices, n_products = engine.PRESETS[precision]
    n = dt_block.shape[1]
    dc_h = dc.detach().cpu().numpy() if isinstance(dc, torch.Tensor) else np.asarray(dc)
    Qt, crank, _ = covariate_basis(dc_h)
    if n <= crank + dimreduce + 1:
        raise ValueError('Insufficient number of cells: must be greater than degrees of freedom '
                         'removed + covariate + 1.')
    Qt_dev = torch.from_numpy(Qt).to(ctx.device) if crank else None
    blk = row_split(n_gene, world)
    local = residualize_block(ctx, dt_block, Qt_dev, n_slices, blk)
    full = gather_sliced(local, n_gene, group) if world > 1 else local
    full.rows = n_gene
    t = (n_gene + TILE - 1) // TILE
    a, b = strip_bounds(t, world)[rank]
    r0, r1 = a * TILE, min(b * TILE, n_gene)
    if out is None:
        P = torch.zeros((max(r1 - r0, 0), n_gene), dtype=torch.float64, device=ctx.device)
        D = torch.zeros_like(P)
    else:
        P, D = out
    if r1 > r0:
        _contract_strip(ctx, MODE_COEX_UPPER, full, full, strip_tiles(t, a, b), (n - 1 - crank - dimreduce) / 2,
                        P, D, r0, n_products)
    return P, D, full.var[:n_gene], (r0, r1)


def coex_host(dt_block_host, dc, n_gene, group=None, precision="default", dimreduce=0, out_dev=None,
              out_host=None):
    """``coex_sharded`` for HOST inputs and outputs: this rank's gene block is a CPU tensor / numpy
    array (pinned memory makes the staged copies asynchronous and overlapped with the projection
    kernels); P and dot strips are copied back into ``out_host`` (CPU tensors) if given.
    Returns (P_strip, dot_strip, var, (row_begin, row_end)) as numpy arrays."""
    from .association import _residualize_any
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ctx = engine.context(None)
    n_slices, n_products = engine.PRESETS[precision]
    xh = dt_block_host if isinstance(dt_block_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dt_block_host))
    n = xh.shape[1]
    dc_h = dc.detach().cpu().numpy() if isinstance(dc, torch.Tensor) else np.asarray(dc)
    Qt, crank, _ = covariate_basis(dc_h)
    if n <= crank + dimreduce + 1:
        raise ValueError('Insufficient number of cells: must be greater than degrees of freedom '
                         'removed + covariate + 1.')
    with torch.cuda.device(ctx.device):
        Qt_dev = torch.from_numpy(Qt).to(ctx.device) if crank else None
        blk = row_split(n_gene, world)
        local = engine.Sliced(blk, n, n_slices, ctx.device)
        if xh.shape[0] < blk:
            local.slices.zero_(); local.quantum.fill_(1.0); local.var.fill_(1.0)
        if xh.shape[0]:
            _residualize_any(ctx, xh, Qt_dev, n_slices, False, out=local, row_offset=0)
        full = gather_sliced(local, n_gene, group) if world > 1 else local
        full.rows = n_gene
        t = (n_gene + TILE - 1) // TILE
        a, b = strip_bounds(t, world)[rank]
        r0, r1 = a * TILE, min(b * TILE, n_gene)
        if out_dev is None:
            P = torch.zeros((max(r1 - r0, 1), n_gene), dtype=torch.float64, device=ctx.device)
            D = torch.zeros_like(P)
        else:
            P, D = out_dev
        if r1 > r0:
            _contract_strip(ctx, MODE_COEX_UPPER, full, full, strip_tiles(t, a, b),
                            (n - 1 - crank - dimreduce) / 2, P, D, r0, n_products)
        if out_host is not None:
            out_host[0].copy_(P, non_blocking=True)
            out_host[1].copy_(D, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            Ph, Dh = out_host[0].numpy(), out_host[1].numpy()
        else:
            Ph, Dh = P.cpu().numpy(), D.cpu().numpy()
        return Ph, Dh, full.var[:n_gene].cpu().numpy(), (r0, r1)


def _contract_strip(ctx, mode, A, B, tiles, dof_a, P, D, row0, n_products, k_chunk=None):
    """Contract with outputs stored from global row ``row0``: hand the C ABI a base pointer that
    is row0 rows before the strip buffers (it only dereferences rows of the listed tiles)."""
    from . import _lib
    ld = D.stride(0)
    tiles = np.ascontiguousarray(tiles, dtype=np.int32)
    off = row0 * ld * 8
    if k_chunk is None:
        k_chunk = engine.plan_k_chunk(A, B, n_products)
    st = ctx.lib.nsr_contract(
        ctx.handle, engine._stream(), engine.ENGINE_UMMA, mode,
        A.slices.data_ptr(), A.rows, A.rows_alloc, A.quantum.data_ptr(), A.var.data_ptr(),
        B.slices.data_ptr(), B.rows, B.rows_alloc, B.quantum.data_ptr(), B.var.data_ptr(),
        A.n, A.n_pad, A.n_slices, n_products, tiles.ctypes.data, tiles.shape[0], float(dof_a),
        P.data_ptr() - off, D.data_ptr() - off, ld, int(k_chunk))
    _lib.check(st, "nsr_contract")
    engine.LAUNCHES += 1 if not k_chunk else -(-A.n_pad // k_chunk)


def gather_dense(P_strip, D_strip, bounds, n_gene, group=None, dst=0):
    """Assemble the full symmetric (n_gene, n_gene) P and dot on rank ``dst`` (None elsewhere)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = P_strip.device
    t = (n_gene + TILE - 1) // TILE
    strips = strip_bounds(t, world)
    outs = []
    for src_t in (P_strip, D_strip):
        full = torch.zeros((n_gene, n_gene), dtype=torch.float64, device=dev) if rank == dst else None
        for k, (a, b) in enumerate(strips):
            r0, r1 = a * TILE, min(b * TILE, n_gene)
            if r1 <= r0:
                continue
            if k == dst:
                if rank == dst:
                    full[r0:r1] = src_t
            elif rank == dst:
                buf = torch.empty((r1 - r0, n_gene), dtype=torch.float64, device=dev)
                dist.recv(buf, src=k, group=group)
                full[r0:r1] = buf
            elif rank == k:
                dist.send(src_t.contiguous(), dst=dst, group=group)
        if rank == dst:
            up = torch.triu(full, 1)
            full = up + up.T
        outs.append(full)
    return outs[0], outs[1]


def de_sharded(dg, dt_block, dc, n_gene, group=None, precision="default", dimreduce=0):
    """DE (single=0) with genes sharded over ranks: no exchange step at all (dg and dc are small
    and replicated).  Returns (P, gamma, varg, vart_block) for this rank's genes."""
    ctx = engine.context(dt_block.device)
    n_slices, n_products = engine.PRESETS[precision]
    n = dt_block.shape[1]
    dc_h = dc.detach().cpu().numpy() if isinstance(dc, torch.Tensor) else np.asarray(dc)
    Qt, crank, _ = covariate_basis(dc_h)
    Qt_dev = torch.from_numpy(Qt).to(ctx.device) if crank else None
    A = engine.residualize(ctx, dg.to(ctx.device, torch.float64), Qt_dev, n_slices)
    B = engine.residualize(ctx, dt_block, Qt_dev, n_slices)
    P = torch.empty((A.rows, B.rows), dtype=torch.float64, device=ctx.device)
    G = torch.empty_like(P)
    engine.contract(ctx, MODE_DE, A, B, engine.rect_tiles(A.rows, B.rows), (n - 1 - crank - dimreduce) / 2,
                    P, G, n_products)
    return P, G, A.var, B.var
