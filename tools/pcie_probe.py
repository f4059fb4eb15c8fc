"""Host<->device copy rates on this box (pinned memory), alone and both directions at once.
Context for the e2e number: the 100k x 20k coex call moves 16 GB in and 6.4 GB out."""
import json
import torch

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
        out["h2d_GBs"] = rate(h2d, nb, [s1])
        out["d2h_GBs"] = rate(d2h, nb, [s2])
        out["duplex_total_GBs"] = rate(both, 2 * nb, [s1, s2])
    print(json.dumps(out))

if __name__ == "__main__":
    main()
