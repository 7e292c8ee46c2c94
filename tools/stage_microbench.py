"""Per-stage micro-benchmark on one 12 MP (or given) frame: CUDA-event timing of each libhhsr stage and the merge
kernel's achieved HBM GB/s.  Used for ncu captures:  ncu --set full -k regex:accumulate_kernel ... python tools/stage_microbench.py
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=3000)
    ap.add_argument("--W", type=int, default=4000)
    ap.add_argument("--scale", type=float, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import bench
    from handheld_super_resolution import alignment as AL, merge as MG, robustness as RB
    from handheld_super_resolution.kernels import estimate_kernels
    from handheld_super_resolution.raw2rgb import postprocess
    from handheld_super_resolution.synthetic import synth_burst
    from handheld_super_resolution.utils_image import compute_grey_images
    scale = int(a.scale) if a.scale == int(a.scale) else a.scale
    burst, _ = synth_burst(5, a.H, a.W, seed=0, device="cuda", as_numpy=False)
    cfg = bench.make_config(scale, a.H, a.W, burst[0].mean().item())
    ref, img = burst[0], burst[1]
    cfa, wb = cfg.exif.cfa_pattern, cfg.exif.white_balance
    std = torch.as_tensor(cfg.noise_model.std_curve, dtype=torch.float64, device="cuda")
    diff = torch.as_tensor(cfg.noise_model.diff_curve, dtype=torch.float64, device="cuda")
    refal = AL.init_alignment(compute_grey_images(ref, "FFT"), cfg)
    rm, rs = RB.init_robustness(ref, cfa, wb, cfg)
    grey = compute_grey_images(img, "FFT")
    flow = AL.align(*refal, grey, cfg)
    table = RB.noise_table((std, diff))      # built once per burst, like main()
    r = RB.compute_robustness(img, rm, rs, flow, cfa, wb, table, cfg)
    covs = estimate_kernels(img, cfg)
    hs, ws = round(scale * a.H), round(scale * a.W)
    num = torch.zeros((hs, ws, 3), device="cuda")
    den = torch.zeros((hs, ws, 3), device="cuda")

    def timeit(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    stages = {
        "grey_fft": lambda: compute_grey_images(img, "FFT"),
        "pyramid": lambda: AL.build_gaussian_pyramid(grey, cfg.block_matching.tuning.factors),
        "align": lambda: AL.align(*refal, grey, cfg),
        "robustness": lambda: RB.compute_robustness(img, rm, rs, flow, cfa, wb, table, cfg),
        "estimate_kernels": lambda: estimate_kernels(img, cfg),
        "merge": lambda: MG.merge(img, flow, covs, r, num, den, cfa, cfg),
        "merge_batch4": lambda: MG.merge_batch([burst[k] for k in (1, 2, 3, 4)], [flow] * 4, [covs] * 4, [r] * 4, num, den, cfa, cfg),
        "merge_ref": lambda: MG.merge_ref(ref, covs, num, den, cfa, cfg),
        "postprocess_u8": lambda: postprocess(None, num, False, False, True, cfg.postprocessing.sharpening, False, None, output_dtype="uint8"),
    }
    res = {}
    for k, fn in stages.items():
        if a.only and k not in a.only.split(","):
            continue
        res[k + "_ms"] = timeit(fn)
    if "merge_ms" in res:
        alg = bench.merge_algorithmic_bytes(a.H, a.W, scale, flow.shape[0], flow.shape[1])
        res["merge_GBps"] = alg / (res["merge_ms"] * 1e-3) / 1e9
        res["merge_frac_of_peak"] = res["merge_GBps"] / bench.measured_peak_gbs()[0]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
