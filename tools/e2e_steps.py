"""Per-step times of the host-to-host legs of bench.py (float32 and uint16 bursts), to see whether a slow average is a
few stalled steps or a uniformly slow pipeline.  Usage: python tools/e2e_steps.py [merge_batch]"""
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution.synthetic import synth_burst
    if len(sys.argv) > 1:
        SR.MERGE_BATCH = int(sys.argv[1])
    wl = bench.WORKLOADS["20x12MP_s2"]
    n, H, W, scale = wl["n"], wl["H"], wl["W"], wl["scale"]
    burst_dev, _ = synth_burst(n, H, W, seed=0, device="cuda", as_numpy=False)
    cfg = bench.make_config(scale, H, W, burst_dev[0].mean().item())
    burst_host = torch.empty((n, H, W), dtype=torch.float32).pin_memory()
    burst_host.copy_(burst_dev)
    import copy
    cfg_u16 = copy.deepcopy(cfg)
    cfg_u16.exif.black_levels, cfg_u16.exif.white_level = [1024, 1024, 1024, 1024], 16383
    wbn = torch.tensor([[cfg.exif.white_balance[c] / cfg.exif.white_balance[1] for c in row] for row in cfg.exif.cfa_pattern],
                       device="cuda", dtype=torch.float32).repeat(H // 2, W // 2)
    counts = torch.round(burst_dev / wbn * (16383 - 1024) + 1024).clamp_(0, 65535).to(torch.int32)
    burst_u16 = torch.empty((n, H, W), dtype=torch.uint16).pin_memory()
    burst_u16.view(torch.int16).copy_(counts.to(torch.int16))
    del burst_dev, counts, wbn
    out_hosts = [torch.empty((scale * H, scale * W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h = torch.cuda.Stream()
    state = {"k": 0, "ev": [None, None]}

    def step(burst, c):
        t0 = time.perf_counter()
        out, _ = SR.main(burst[0], burst[1:], c)
        t1 = time.perf_counter()
        k = state["k"] % 2
        state["k"] += 1
        if state["ev"][k] is not None:
            state["ev"][k].synchronize()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(d2h):
            d2h.wait_event(ready)
            out_hosts[k].copy_(out, non_blocking=True)
            out.record_stream(d2h)
            done = torch.cuda.Event(enable_timing=True)
            done.record()
        state["ev"][k] = done
        return done, (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3

    def lat_leg():      # bench.py's single-burst-latency leg: one burst at a time, D2H on the compute stream, host-synchronous
        for _ in range(7):
            out, _ = SR.main(burst_host[0], burst_host[1:], cfg)
            out_hosts[0].copy_(out, non_blocking=True)
            torch.cuda.synchronize()

    legs = [("float32", burst_host, cfg), ("uint16", burst_u16, cfg_u16), ("float32 again", burst_host, cfg), ("uint16 again", burst_u16, cfg_u16)]
    if os.environ.get("LAT_LEG"):
        legs = [("float32", burst_host, cfg), ("LAT", None, None), ("uint16", burst_u16, cfg_u16), ("LAT", None, None),
                ("uint16 again", burst_u16, cfg_u16), ("float32 again", burst_host, cfg), ("LAT", None, None), ("uint16 3", burst_u16, cfg_u16)]
    for name, burst, c in legs:
        if name == "LAT":
            lat_leg()
            continue
        for _ in range(2):
            step(burst, c)
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        nsteps = int(os.environ.get("STEPS", "12"))
        a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        marks, mallocs = [], []
        for _ in range(nsteps):
            marks.append(step(burst, c))
            mallocs.append(torch.cuda.memory_stats().get("num_device_alloc", 0) - a0)
        torch.cuda.synchronize()
        ends = [start.elapsed_time(m[0]) for m in marks]
        per = [round(b - a, 2) for a, b in zip([0.0] + ends[:-1], ends)]
        print(name, "per-step ms (result landed on the host):", per)
        print("   cudaMalloc calls so far:", mallocs)
        print("   host enqueue ms:", [round(m[1], 1) for m in marks], " wait-for-buffer ms:", [round(m[2], 1) for m in marks])


if __name__ == "__main__":
    main()
