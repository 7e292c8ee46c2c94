"""Run under torchrun (N >= 2): main_sharded() in every exchange mode against single-process main() on the same burst.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py

Prints one JSON line per mode (rank 0) and exits non-zero on a mismatch:
  rows            bit-identical to the single-GPU image (same kernel, frames accumulated in burst order), slice by slice;
  p2p / reduce_scatter / allreduce   float32 summation order only (< 1e-5).
tests/test_gpu_multi.py runs this when the box has at least 2 GPUs."""
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
    from handheld_super_resolution import main as sr_main
    from handheld_super_resolution.distributed import main_sharded
    from handheld_super_resolution.synthetic import synth_burst
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    cases = [(1504, 2016, 9, 2, "steerable"), (1000, 1400, 4, 3, "steerable"), (640, 704, 2, 2, "iso")]
    for H, W, n, scale, kern in cases:
        burst, _ = synth_burst(n, H, W, seed=3, device="cuda", as_numpy=False)
        cfg = bench.make_config(scale, H, W, burst[0].mean().item())
        cfg.merging.kernel = kern
        if min(H, W) < 673:
            cfg.block_matching.tuning.factors = [1, 2, 2, 2]
        out_1, dbg_1 = sr_main(burst[0], burst[1:], cfg)
        for mode in ("rows", "p2p", "reduce_scatter", "allreduce"):
            for rep in range(2):                       # twice: buffers of the first burst are reused by the second
                out_s, dbg_s = main_sharded(burst[0], burst[1:], cfg, mode=mode)
                torch.cuda.synchronize()
            rows = dbg_s.get("rows")
            res = torch.zeros(3, device="cuda")
            if mode == "rows":
                want = out_1[rows[0]:rows[1]]
                res[0] = torch.nan_to_num((out_s - want).abs(), nan=0.0).max() if want.numel() else 0.0
                res[1] = 0.0 if torch.equal(torch.isnan(out_s), torch.isnan(want)) else 1.0
                lr0, lr1 = int(rows[0] / scale), min(H, int(-(-rows[1] // scale)) + 1)
                res[2] = (dbg_s["accumulated robustness"][lr0:lr1] - dbg_1["accumulated robustness"][lr0:lr1]).abs().max()
            elif rank == 0 or mode != "p2p":
                res[0] = torch.nan_to_num((out_s - out_1).abs(), nan=0.0).max()
                res[1] = 0.0 if torch.equal(torch.isnan(out_s), torch.isnan(out_1)) else 1.0
                res[2] = (dbg_s["accumulated robustness"] - dbg_1["accumulated robustness"]).abs().max()
            dist.all_reduce(res, op=dist.ReduceOp.MAX)
            tol = 0.0 if mode == "rows" else 1e-5
            good = res[0].item() <= tol and res[1].item() == 0 and res[2].item() < 1e-9
            ok = ok and good
            if rank == 0:
                print(json.dumps({"world": world, "case": [H, W, n, scale, kern], "mode": mode, "max_abs_diff": res[0].item(),
                                  "tolerance": tol, "nan_pattern_differs": bool(res[1].item()), "acc_rob_diff": res[2].item(),
                                  "ok": good}), flush=True)
        del burst, out_1
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
