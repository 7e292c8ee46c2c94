"""Measurement tooling (NOT framework code): wall time of the UNMODIFIED reference (baseline/_ref, its own Numba-CUDA +
torch path) on one benchmark workload on this box's GPU — the number that sits beside this repository's in the bench
line (`bench.py --impl reference` embeds it as `reference_numba_gpu`).  Same synthetic burst and configuration as the
CUDA arm of bench.py (ISO-100 curves of the reference, SNR as process() derives it, tile size 32).

    python baseline/time_reference_numba.py 20 3000 4000 2      # prints one JSON line
"""
import contextlib
import io
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    n, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    scale = float(sys.argv[4])
    scale = int(scale) if scale == int(scale) else scale
    import probe_reference as P          # stubs the absent third-party modules, puts baseline/_ref on sys.path
    import importlib.util
    import numpy as np
    import torch
    spec = importlib.util.spec_from_file_location(
        "hhsr_synthetic", os.path.join(HERE, "..", "handheld-multi-frame-super-resolution_b200", "handheld_super_resolution", "synthetic.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    P.install_copysign_shim()
    from numba import cuda
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution.params import sanitize_config
    burst_t, _ = synth.synth_burst(n, H, W, seed=0, device="cuda", as_numpy=False)
    burst = burst_t.cpu().numpy()
    del burst_t
    torch.cuda.empty_cache()
    cfg = P.make_config(scale, 32, burst[0])
    cfg.verbose = 0
    sanitize_config(cfg, burst[0].shape)
    times = []
    for rep in range(3):                  # the first run compiles the kernels
        cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out, _ = SR.main(burst[0], burst[1:], cfg)
        cuda.synchronize()
        times.append(time.perf_counter() - t0)
        del out
    hs, ws = round(scale * H), round(scale * W)
    best = min(times[1:])
    print(json.dumps({"impl": "the unmodified reference, Numba-CUDA on this GPU (host burst in, device image out, verbose 0)",
                      "gpu": torch.cuda.get_device_name(0), "workload": [n, H, W, scale], "first_run_with_jit_s": times[0],
                      "warm_runs_s": times[1:], "ms_per_step": best * 1e3, "value": hs * ws / 1e6 / best, "unit": "MPix/s"}))


if __name__ == "__main__":
    main()
