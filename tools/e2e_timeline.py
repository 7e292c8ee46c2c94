"""Timeline of the e2e step of bench.py (host burst -> main() -> host image) from torch.profiler (CUPTI): per stream,
busy time of kernels / H2D / D2H copies, and the gaps, to see what bounds the end-to-end number."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    from handheld_super_resolution.distributed import main_sharded
    from handheld_super_resolution.synthetic import synth_burst
    from torch.profiler import ProfilerActivity, profile
    wl = bench.WORKLOADS["20x12MP_s2"]
    n, H, W, scale = wl["n"], wl["H"], wl["W"], wl["scale"]
    burst_dev, _ = synth_burst(n, H, W, seed=0, device="cuda", as_numpy=False)
    cfg = bench.make_config(scale, H, W, burst_dev[0].mean().item())
    burst_host = torch.empty((n, H, W), dtype=torch.float32).pin_memory()
    burst_host.copy_(burst_dev)
    del burst_dev
    out_hosts = [torch.empty((scale * H, scale * W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h = torch.cuda.Stream()
    state = {"k": 0, "ev": [None, None]}

    def step():
        out, _ = main_sharded(burst_host[0], burst_host[1:], cfg)
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
            done = torch.cuda.Event()
            done.record()
        state["ev"][k] = done

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in evs)
    t1 = max(e.time_range.end for e in evs)
    rows = {}
    for e in evs:
        kind = "H2D" if "HtoD" in e.name else "D2H" if "DtoH" in e.name else "memset" if "emset" in e.name else "kernel"
        r = rows.setdefault(kind, {"n": 0, "busy_us": 0.0, "first": 1e30, "last": 0.0})
        r["n"] += 1
        r["busy_us"] += e.time_range.end - e.time_range.start
        r["first"] = min(r["first"], e.time_range.start - t0)
        r["last"] = max(r["last"], e.time_range.end - t0)
    print(json.dumps({"span_ms_3_steps": (t1 - t0) / 1e3, "per_kind": rows}, indent=1))
    # per-copy details of the H2D copies of one step: start offsets and durations
    h2d = sorted([e for e in evs if "HtoD" in e.name and (e.time_range.end - e.time_range.start) > 200], key=lambda e: e.time_range.start)
    print("H2D copies > 200us: n=%d" % len(h2d))
    for e in h2d[:24]:
        print("  start %.3f ms  dur %.3f ms" % ((e.time_range.start - t0) / 1e3, (e.time_range.end - e.time_range.start) / 1e3))
    d2hs = sorted([e for e in evs if "DtoH" in e.name and (e.time_range.end - e.time_range.start) > 200], key=lambda e: e.time_range.start)
    for e in d2hs:
        print("  D2H start %.3f ms  dur %.3f ms" % ((e.time_range.start - t0) / 1e3, (e.time_range.end - e.time_range.start) / 1e3))
    # the long CPU-side ops (cudaMalloc, cudaFree, synchronisations)
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and
           any(s in e.name for s in ("cudaMalloc", "cudaFree", "Synchronize", "cudaHostAlloc"))]
    agg = {}
    for e in cpu:
        a = agg.setdefault(e.name, [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    print({k: (v[0], round(v[1] / 1e3, 2)) for k, v in agg.items()})


if __name__ == "__main__":
    main()
