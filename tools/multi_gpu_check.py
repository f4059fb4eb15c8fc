#!/usr/bin/env python3
"""Run under torchrun (N >= 2 GPUs): the sharded co-expression must reproduce the single-GPU
result bit for bit (integer sums are order independent; the epilogue is the same code).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [genes] [cells]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from normalisr_b200 import association, engine, parallel, synth  # noqa: E402


def main():
    genes = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p = synth.device_problem(77, genes, cells, dev)          # same seed -> same full matrix on every rank
    dt, dc = p["dt"], p["dc"]
    blk = parallel.row_split(genes, world)
    mine = dt[rank * blk:(rank + 1) * blk]
    ok = same = True
    if rank == 0:
        res = association.association_tests(dt, None, dc)
        P1, D1, var1 = res[0], res[1], res[4]
    for schedule, transport in (("pairs", "nccl"), ("pairs", "ce"), ("pairs", "ce"), ("allgather", "nccl")):
        parallel.TRANSPORT = transport
        P_s, D_s, var, (r0, r1) = parallel.coex_sharded(mine, dc, genes, schedule=schedule)
        P, D = parallel.gather_dense(P_s, D_s, None, genes, schedule=schedule)
        if rank == 0:
            good = bool(torch.equal(P, P1) and torch.equal(D, D1) and torch.equal(var, var1))
            ok = ok and good
            print("multi-GPU check [%s/%s]: world %d genes %d cells %d: identical to single GPU: %s (max|dP| %.3e)" % (
                schedule, transport, world, genes, cells, good, float((P - P1).abs().max())), flush=True)
        # host-buffer API
        Ph, Dh, varh, (a0, a1) = parallel.coex_host(mine.cpu(), dc.cpu().numpy(), genes, schedule=schedule)
        s_ok = bool(torch.equal(torch.from_numpy(Ph).to(dev), P_s) and torch.equal(torch.from_numpy(Dh).to(dev), D_s)
                    and (a0, a1) == (r0, r1))
        same = same and s_ok
        print("rank %d [%s]: rows [%d,%d) host-API output identical to device-API output: %s" % (
            rank, schedule, r0, r1, s_ok), flush=True)
    # the full symmetric matrices, both triangles, written by all ranks into ONE shared page-locked mapping
    parallel.TRANSPORT = "auto"
    (Ph, Dh), shared = parallel.shared_host_matrices(2, (genes, genes))
    for rep in range(2):
        Ph.fill_(-7.0)
        Dh.fill_(-7.0)
        dist.barrier()
        parallel.coex_host(mine.cpu().pin_memory(), dc.cpu().numpy(), genes, home=(Ph, Dh))
        dist.barrier()
        if rank == 0:
            good = (bool(torch.equal(Ph, P1.cpu()) and torch.equal(Dh, D1.cpu())) if shared else
                    bool((Ph == -7.0).any()))       # private fallback: rank 0 holds only its own rectangles
            ok = ok and good
            print("multi-GPU check [home, shared mapping %s, rep %d]: rank 0 holds the complete symmetric P and dot, "
                  "identical to single GPU: %s" % (shared, rep, good), flush=True)
        dist.barrier()
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("coex_host strips identical on all ranks: %s" % bool(flag.item() == 1.0), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
