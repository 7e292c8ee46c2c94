"""Survey-time probe (measurement tooling, NOT framework code).

Runs the *unmodified* reference hot path (baseline/_ref/handheld_super_resolution, a verbatim copy of
/root/reference/handheld_super_resolution) on a real GPU through its own Numba-CUDA + torch path, on a synthetic
Bayer burst, bypassing only DNG I/O (process() -> main()).  Third-party modules that the non-hot-path files import
and that are absent from this image (omegaconf, rawpy, exifread, imageio, skimage, matplotlib) are stubbed.

Usage (GPU box):  python baseline/probe_reference.py [n_frames H W scale Ts]
Writes gpurun_out/probe_reference.json (+ .log).
"""
import json
import os
import sys
import time
import traceback
from unittest import mock

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
os.makedirs(OUT, exist_ok=True)
RES = {}


def dump():
    with open(os.path.join(OUT, "probe_reference.json"), "w") as f:
        json.dump(RES, f, indent=1, default=str)


for name in ["omegaconf", "rawpy", "exifread", "imageio", "skimage", "skimage.filters", "matplotlib",
             "matplotlib.pyplot"]:
    sys.modules[name] = mock.MagicMock()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


class Cfg(dict):
    """Minimal attribute-dict standing in for an OmegaConf DictConfig (attribute + item access, .get, .update)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return Cfg({k: Cfg.wrap(v) for k, v in d.items()})
        return d


def make_config(scale, Ts, ref_frame):
    import yaml
    c = Cfg.wrap(yaml.safe_load(open(os.path.join(REF, "configs", "default.yaml"))))
    c.scale = scale
    c.verbose = 2
    c.block_matching.tuning.tile_size = Ts
    c.noise_model.alpha = 1.80710882e-4
    c.noise_model.beta = 3.1937599182128e-6
    from handheld_super_resolution.params import update_snr_config
    c.exif = Cfg(cfa_pattern=[[0, 1], [1, 2]], iso=100, white_balance=[2.0, 1.0, 1.5, 0.0])
    c.noise_model.std_curve = np.load(os.path.join(REF, "data", "noise_model_std_ISO_100.npy")).tolist()
    c.noise_model.diff_curve = np.load(os.path.join(REF, "data", "noise_model_diff_ISO_100.npy")).tolist()
    c.accumulated_robustness_denoiser.enabled = False
    # SNR exactly as process() derives it (super_resolution.py:261-274)
    brightness = float(np.mean(ref_frame))
    snr = brightness / c.noise_model.std_curve[round(1000 * brightness)]
    update_snr_config(c, snr)
    return c


def synth_burst(n, H, W, seed=0, max_shift=3.0):
    """Band-limited random RGB scene at 2x, per-frame sub-pixel translation, 2x box decimation, RGGB mosaic,
    heteroscedastic Gaussian noise (alpha*I+beta).  Returns float32 [n,H,W] in [0,1] and the true (dy,dx) shifts."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    up = 2
    pad = 16
    hh, ww = H * up + 2 * pad, W * up + 2 * pad
    scene = torch.rand((1, 3, hh // 4 + 2, ww // 4 + 2), device="cuda", generator=g)
    scene = torch.nn.functional.interpolate(scene, size=(hh, ww), mode="bicubic", align_corners=False)
    fine = torch.rand((1, 3, hh, ww), device="cuda", generator=g)
    k = torch.tensor([1, 4, 6, 4, 1.0], device="cuda") / 16
    fine = torch.nn.functional.conv2d(fine, k.view(1, 1, 5, 1).repeat(3, 1, 1, 1), groups=3, padding=(2, 0))
    fine = torch.nn.functional.conv2d(fine, k.view(1, 1, 1, 5).repeat(3, 1, 1, 1), groups=3, padding=(0, 2))
    scene = (0.6 * scene + 0.4 * fine).clamp(0, 1) * 0.8 + 0.05
    rng = np.random.default_rng(seed)
    frames, shifts = [], []
    ys, xs = torch.meshgrid(torch.arange(H * up, device="cuda", dtype=torch.float32),
                            torch.arange(W * up, device="cuda", dtype=torch.float32), indexing="ij")
    for i in range(n):
        dy, dx = (0.0, 0.0) if i == 0 else rng.uniform(-max_shift, max_shift, 2)
        shifts.append((float(dy), float(dx)))
        gy = (ys + pad - dy * up + 0.5) / hh * 2 - 1  # frame[y] = scene[y - dy]  =>  reference flow == +(dx, dy)
        gx = (xs + pad - dx * up + 0.5) / ww * 2 - 1
        grid = torch.stack([gx, gy], -1)[None]
        sh = torch.nn.functional.grid_sample(scene, grid, mode="bilinear", align_corners=False)
        lr = torch.nn.functional.avg_pool2d(sh, up)[0]  # [3,H,W]
        bay = torch.empty((H, W), device="cuda")
        bay[0::2, 0::2] = lr[0, 0::2, 0::2]
        bay[0::2, 1::2] = lr[1, 0::2, 1::2]
        bay[1::2, 0::2] = lr[1, 1::2, 0::2]
        bay[1::2, 1::2] = lr[2, 1::2, 1::2]
        noise = torch.randn((H, W), device="cuda", generator=g)
        bay = (bay + torch.sqrt(1.80710882e-4 * bay + 3.1937599182128e-6) * noise).clamp(0, 1)
        frames.append(bay.cpu().numpy().astype(np.float32))
    del scene, fine
    torch.cuda.empty_cache()
    return np.stack(frames), shifts


def timed_main(SR, burst, cfg, label):
    import io
    import contextlib
    from numba import cuda
    buf = io.StringIO()
    cuda.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        out, dbg = SR.main(burst[0], burst[1:], cfg)
    cuda.synchronize()
    dt = time.perf_counter() - t0
    log = buf.getvalue()
    with open(os.path.join(OUT, "probe_reference.log"), "a") as f:
        f.write("\n===== %s =====\n" % label + log)
    stage = {}
    for line in log.splitlines():
        if ":" in line and "milliseconds" in line:
            k, v = line.rsplit(":", 1)
            k = k.strip()
            ms = float(v.replace("milliseconds", "").strip())
            stage.setdefault(k, []).append(ms)
    RES[label] = {"wall_s": dt, "stage_ms_sum": {k: round(sum(v), 2) for k, v in stage.items()},
                  "stage_ms_n": {k: len(v) for k, v in stage.items()}}
    dump()
    return out, dbg


def install_copysign_shim():
    """numba 0.65 + NVVM 12.9 reject libdevice's __nv_copysign ("Unsupported intrinsic: llvm.copysign.f64"): exact
    bit-twiddling lowering of math.copysign(f64, f64) instead (same shim as tests/golden/make_golden_gpu.py)."""
    import math
    from llvmlite import ir
    from numba import types
    from numba.cuda.mathimpl import lower

    @lower(math.copysign, types.float64, types.float64)
    def copysign_f64(context, builder, sig, args):
        i64 = ir.IntType(64)
        xi, yi = builder.bitcast(args[0], i64), builder.bitcast(args[1], i64)
        mag = builder.and_(xi, ir.Constant(i64, 0x7FFFFFFFFFFFFFFF))
        sgn = builder.and_(yi, ir.Constant(i64, 0x8000000000000000))
        return builder.bitcast(builder.or_(mag, sgn), ir.DoubleType())


def main():
    install_copysign_shim()
    n, H, W, scale, Ts = 8, 3000, 4000, 2, 32
    if len(sys.argv) > 5:
        n, H, W, scale, Ts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
        scale = int(scale) if scale == int(scale) else scale
    RES["env"] = {"torch": torch.__version__, "cuda": torch.version.cuda, "gpu": torch.cuda.get_device_name(0),
                  "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "nproc": os.cpu_count()}
    import numba
    from numba import cuda
    RES["env"]["numba"] = numba.__version__
    RES["env"]["numba_cuda_available"] = cuda.is_available()
    RES["env"]["numba_cc"] = str(cuda.get_current_device().compute_capability)
    dump()
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution import block_matching as BM
    from handheld_super_resolution.params import sanitize_config

    # ---- 1. L1 local search behaviour (is it round-only for Ts 32/64? what does Ts 16 do?)
    try:
        for ts, kern in [(16, BM.cuda_L1_local_search16), (32, BM.cuda_L1_local_search32),
                         (64, BM.cuda_L1_local_search64)]:
            g = torch.Generator(device="cuda").manual_seed(1)
            ref = torch.rand((ts * 6, ts * 8), device="cuda", generator=g)
            mov = torch.roll(ref, shifts=(1, -1), dims=(0, 1)).contiguous()  # true shift: dy=+1, dx=-1
            al = (torch.rand((6, 8, 2), device="cuda", generator=g) - 0.5) * 0.8  # |flow|<0.4 -> round() == 0
            al0 = al.clone()
            tpb = (ts, ts) if ts < 64 else (64, 16)
            reps = []
            for _ in range(3):
                a = al0.clone()
                kern[(8, 6), tpb](ref, mov, 1, a)
                cuda.synchronize()
                reps.append(a.cpu().numpy())
            RES["L1_ts%d" % ts] = {
                "equals_round_of_input": bool(np.array_equal(reps[0], np.round(al0.cpu().numpy()))),
                "deterministic_3_runs": bool(all(np.array_equal(reps[0], r) for r in reps)),
                "unique_dx": np.unique(reps[0][..., 0]).tolist(), "unique_dy": np.unique(reps[0][..., 1]).tolist(),
                "true_shift_dx_dy": [-1, 1]}
            dump()
    except Exception:
        RES["L1_error"] = traceback.format_exc()
        dump()

    # ---- 2. end-to-end reference main() on the synthetic burst
    try:
        t0 = time.perf_counter()
        burst, shifts = synth_burst(n, H, W)
        RES["synth"] = {"n": n, "H": H, "W": W, "scale": scale, "Ts": Ts, "gen_s": time.perf_counter() - t0,
                        "shifts_dy_dx": shifts, "mean": float(burst.mean())}
        cfg = make_config(scale, Ts, burst[0])
        sanitize_config(cfg, burst[0].shape)
        RES["synth"]["tile_sizes"] = list(cfg.block_matching.tuning.tile_sizes)
        flows = []
        _al = SR.align

        def cap(*a, **k):
            r = _al(*a, **k)
            flows.append(r.detach().clone())
            return r
        SR.align = cap
        out, _ = timed_main(SR, burst, cfg, "run1_with_jit")
        fl = torch.stack(flows).cpu().numpy()
        med = np.median(fl.reshape(fl.shape[0], -1, 2), axis=1)
        RES["flow_check"] = {"median_flow_dx_dy": med.tolist(),
                             "true_dx_dy": [[s[1], s[0]] for s in shifts[1:]],
                             "frac_integer_valued": float(np.mean(fl == np.round(fl)))}
        o = out.copy_to_host()
        RES["out"] = {"shape": list(o.shape), "dtype": str(o.dtype), "nan": int(np.isnan(o).sum()),
                      "min": float(np.nanmin(o)), "max": float(np.nanmax(o)), "mean": float(np.nanmean(o))}
        dump()
        del out, o
        flows.clear()
        for rep in range(2):
            timed_main(SR, burst, cfg, "run%d_warm" % (rep + 2))
            flows.clear()
        # warm, verbose=0 (no per-stage synchronize): the number comparable to the README's "<4 s on RTX 3090"
        cfg.verbose = 0
        for rep in range(2):
            timed_main(SR, burst, cfg, "run%d_warm_quiet" % (rep + 4))
            flows.clear()
        # same with TF32 convs disabled (clean-fp32 oracle mode)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        out2, _ = timed_main(SR, burst, cfg, "run6_warm_quiet_notf32")
        RES["peak_mem_GB_torch"] = torch.cuda.max_memory_allocated() / 1e9
        free, total = torch.cuda.mem_get_info()
        RES["mem_used_GB_device"] = (total - free) / 1e9
        dump()
    except Exception:
        RES["main_error"] = traceback.format_exc()
        dump()
    print(json.dumps(RES, indent=1, default=str)[:6000])


if __name__ == "__main__":
    main()
