"""Host<->device copy rates on this box (pinned memory), alone and both directions at once.
Context for the e2e number: the 100k x 20k coex call moves 16 GB in and 6.4 GB out."""
import json
import os
import torch
import torch.distributed as dist

def rate(fn, nbytes, streams):
    for s in streams:
        s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # under torchrun: every rank copies at the same time (barrier before each measurement); rank 0 prints the
        # per-rank rates and the aggregate = total bytes / slowest rank's time
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    gb = 4
    h_in = torch.empty(gb << 27, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(gb << 27, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(gb << 27, dtype=torch.float64, device="cuda")
    d_out = torch.ones(gb << 27, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    nb = gb << 30
    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    def both():
        h2d(); d2h()
    out = {}
    for _ in range(2):
        for key, fn, n, ss in (("h2d_GBs", h2d, nb, [s1]), ("d2h_GBs", d2h, nb, [s2]),
                               ("duplex_total_GBs", both, 2 * nb, [s1, s2])):
            if world > 1:
                dist.barrier()
            out[key] = rate(fn, n, ss)
    if world > 1:
        t = torch.tensor([out["h2d_GBs"], out["d2h_GBs"], out["duplex_total_GBs"]], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        if dist.get_rank() == 0:
            m = torch.stack(allr).cpu()
            agg = {k: float(world * m[:, i].min()) for i, k in enumerate(("h2d_GBs", "d2h_GBs", "duplex_total_GBs"))}
            print(json.dumps({"ranks": world, "concurrent": True, "per_rank": m.tolist(),
                              "aggregate_GBs_at_slowest_rank": agg}))
        dist.destroy_process_group()
        return
    print(json.dumps(out))

if __name__ == "__main__":
    main()
