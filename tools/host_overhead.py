"""Host-side cost of one burst through main(): wall time the Python thread spends ENQUEUEING the work (no synchronisation
inside the call) against the GPU time of the same burst.  When enqueue time >= GPU time the pipeline is launch-bound
(what happens to a rank of an 8-GPU run, whose 2-3 frames take ~0.7 ms of GPU time each)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    from handheld_super_resolution import _lib, super_resolution as SR
    from handheld_super_resolution.synthetic import synth_burst
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    burst, _ = synth_burst(n, 3000, 4000, seed=0, device="cuda", as_numpy=False)
    cfg = bench.make_config(2, 3000, 4000, burst[0].mean().item())
    for _ in range(3):
        SR.main(burst[0], burst[1:], cfg)
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = _lib.launch_count
        t0 = time.perf_counter()
        e0.record()
        SR.main(burst[0], burst[1:], cfg)
        e1.record()
        host.append((time.perf_counter() - t0) * 1e3)
        launches = _lib.launch_count - l0
        torch.cuda.synchronize()
        gpu.append(e0.elapsed_time(e1))
    print(json.dumps({"frames": n, "host_enqueue_ms": sorted(host)[len(host) // 2], "gpu_ms": sorted(gpu)[len(gpu) // 2],
                      "libhhsr_launches": launches, "host_us_per_comp_frame": sorted(host)[len(host) // 2] * 1e3 / max(n - 1, 1)}))


if __name__ == "__main__":
    main()
