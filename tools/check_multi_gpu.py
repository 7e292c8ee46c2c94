"""Run under torchrun (N >= 2): frame-sharded main_sharded() against single-process main() on the same burst.
Prints max |difference| (float32 summation order only) and the time of the one NCCL sum."""
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
    H, W, n = 1504, 2016, 9
    cfg = bench.make_config(2, H, W)
    burst, _ = synth_burst(n, H, W, seed=3, device="cuda", as_numpy=False)
    out_s, dbg_s = main_sharded(burst[0], burst[1:], cfg)
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        out_1, dbg_1 = sr_main(burst[0], burst[1:], cfg)
        a, b = out_s, out_1
        same_nan = bool(torch.equal(torch.isnan(a), torch.isnan(b)))
        d = (torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max().item()
        dr = (dbg_s["accumulated robustness"] - dbg_1["accumulated robustness"]).abs().max().item()
        print("world=%d  max|sharded - single| = %.3g (tolerance 1e-5)  same NaN set: %s  acc_rob diff %.3g"
              % (dist.get_world_size(), d, same_nan, dr), flush=True)
        assert d < 1e-5 and same_nan and dr < 1e-9
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
