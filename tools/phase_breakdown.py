"""Phase breakdown of main_sharded() on the benchmark burst (run under torchrun for N > 1): CUDA-event times of the
reference-side products, the frame loop, the one reduction, merge_ref and the re-assembly, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution.distributed import main_sharded
    from handheld_super_resolution.synthetic import synth_burst
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "20x12MP_s2"]
    cfg = bench.make_config(wl["scale"], wl["H"], wl["W"])
    burst, _ = synth_burst(wl["n"], wl["H"], wl["W"], seed=0, device="cuda", as_numpy=False)
    for _ in range(3):
        main_sharded(burst[0], burst[1:], cfg)
    acc = {}
    reps = 5
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        SR.PHASE_EVENTS = []
        main_sharded(burst[0], burst[1:], cfg)
        torch.cuda.synchronize()
        ev = SR.PHASE_EVENTS
        SR.PHASE_EVENTS = None
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / reps
        acc["total"] = acc.get("total", 0.0) + ev[0][1].elapsed_time(ev[-1][1]) / reps
    t = torch.tensor([acc[k] for k in sorted(acc)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"world": world, "phase_ms_max_over_ranks": dict(zip(sorted(acc), [round(x, 3) for x in t.tolist()]))}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
