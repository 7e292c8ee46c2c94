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
    burst, _ = synth_burst(wl["n"], wl["H"], wl["W"], seed=0, device="cuda", as_numpy=False)
    cfg = bench.make_config(wl["scale"], wl["H"], wl["W"], burst[0].mean().item())
    if len(sys.argv) > 2:
        os.environ["HHSR_SHARD_REDUCE"] = sys.argv[2]
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
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank = [t]
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"world": world, "mode": os.environ.get("HHSR_SHARD_REDUCE", "reduce_scatter"),
                          "phase_ms_max_over_ranks": dict(zip(sorted(acc), [round(x, 3) for x in t.tolist()])),
                          "phase_ms_per_rank": {k: [round(p[i].item(), 3) for p in per_rank] for i, k in enumerate(sorted(acc))}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
