"""Host<->device copy bandwidth of the box (pinned memory), alone and with both directions in flight: the e2e number
of bench.py is bounded by these (960 MB of float32 RAW up, 576 MB of float32 image down per 20x12MP burst)."""
import json

import torch


def bw(fn, nbytes, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    n = 256 << 20
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {"h2d_GBps": bw(lambda: d_up.copy_(h_up, non_blocking=True), n),
           "d2h_GBps": bw(lambda: h_dn.copy_(d_dn, non_blocking=True), n)}

    def both():
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
    res["bidir_each_GBps"] = bw(both, n)
    small = 48_000_000
    res["h2d_48MB_GBps"] = bw(lambda: d_up[:small].copy_(h_up[:small], non_blocking=True), small, 20)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
