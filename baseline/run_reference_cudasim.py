"""CPU baseline harness (measurement tooling, NOT framework code): runs the reference pipeline's main() under
NUMBA_ENABLE_CUDASIM=1 on the host cores, on a synthetic Bayer burst.

The reference cannot run under the simulator as shipped.  This harness
  * copies /root/reference/handheld_super_resolution to a scratch dir and applies 4 Python-semantics patches that do
    not change what the compiled Numba-CUDA path computes (1/0 -> inf, undefined min_shift_* -> 0, round(inf) guard,
    shape[-1] -> shape[2]);
  * stubs the third-party modules missing from this image (omegaconf, rawpy, exifread, imageio, skimage, matplotlib);
  * adds the simulator features the reference needs (cuda.as_cuda_array, cuda.shfl_down_sync, torch tensors as kernel
    arguments) and maps torch device "cuda" to "cpu" when no GPU is present.

Verified in the survey container (no GPU, 8 cores): 2 frames 64x64, scale 2, Ts 16, factors [1,2,2] -> 244 s.
The simulator runs one Python thread per CUDA thread (GIL-bound: ~1 core busy), so sizes must stay tiny.

Usage: python baseline/run_reference_cudasim.py --n 2 --size 256 --scale 1 --ts 32 --factors 1,2,2,2 --out out.npz
"""
import argparse
import os
import shutil
import sys
import tempfile
import threading
import time
from unittest import mock

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

REFERENCE = os.environ.get("HHSR_REFERENCE", "/root/reference")


def patched_reference_copy():
    dst = tempfile.mkdtemp(prefix="hhsr_ref_sim_")
    pkg = os.path.join(dst, "handheld_super_resolution")
    shutil.copytree(os.path.join(REFERENCE, "handheld_super_resolution"), pkg)
    os.chmod(pkg, 0o755)
    for f in os.listdir(pkg):
        os.chmod(os.path.join(pkg, f), 0o644)

    def sub(fname, fn):
        p = os.path.join(pkg, fname)
        s = open(p).read()
        s2 = fn(s)
        assert s2 != s, "patch did not apply to " + fname
        open(p, "w").write(s2)

    # robustness.py:390,582-585,678  `1/0` is +inf on the GPU (numpy error model) but ZeroDivisionError in Python
    sub("robustness.py", lambda s: s.replace("-1/0", "-math.inf").replace("+1/0", "math.inf")
        .replace("= 1/0 #", "= math.inf #"))
    # robustness.py:519  round(inf) raises OverflowError in Python; on the GPU it reads a garbage curve entry and the
    # pixel ends with R = 0 either way (d_sq is NaN -> clamp(NaN) = 0)
    sub("robustness.py", lambda s: s.replace(
        "id_noise = round(1000 *brightness)", "id_noise = round(1000 *brightness) if math.isfinite(brightness) else 0"))
    # block_matching.py:179,251,344  min_shift_x/y may be unbound; compiled Numba reads them as 0
    sub("block_matching.py", lambda s: s.replace(
        "    # Now find the minimum error and corresponding shift\n",
        "    min_shift_x = 0\n    min_shift_y = 0\n    # Now find the minimum error and corresponding shift\n"))
    # utils.py:75  the simulator's FakeShape rejects negative indices
    sub("utils.py", lambda s: s.replace("num.shape[-1]", "num.shape[2]"))
    return dst


def install_shims():
    for name in ["omegaconf", "rawpy", "exifread", "imageio", "skimage", "skimage.filters", "matplotlib",
                 "matplotlib.pyplot"]:
        sys.modules[name] = mock.MagicMock()
    if not torch.cuda.is_available():
        def strip(fn):
            def w(*a, **k):
                if k.get("device", None) == "cuda":
                    k["device"] = "cpu"
                return fn(*a, **k)
            return w
        torch.as_tensor = strip(torch.as_tensor)
        torch.zeros = strip(torch.zeros)
    from numba import cuda
    from numba.cuda.simulator import kernelapi, kernel as simkernel
    from numba.cuda.simulator.cudadrv.devicearray import FakeCUDAArray

    cuda.as_cuda_array = lambda t: FakeCUDAArray(t.detach().cpu().numpy())

    def shfl_down_sync(self, mask, value, delta):
        th = threading.current_thread()
        buf = th._manager.__dict__.setdefault("_shfl_buf", {})
        tid = th.thread_id
        buf[tid] = value
        th.syncthreads()
        res = buf[tid + delta] if (tid % 32 + delta < 32 and (tid + delta) in buf) else value
        th.syncthreads()
        return res
    kernelapi.FakeCUDAModule.shfl_down_sync = shfl_down_sync

    orig = simkernel.FakeCUDAKernel.__call__

    def call(self, *args):
        args = [FakeCUDAArray(a.detach().cpu().numpy()) if isinstance(a, torch.Tensor) else a for a in args]
        return orig(self, *args)
    simkernel.FakeCUDAKernel.__call__ = call


class Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        return Cfg({k: Cfg.wrap(v) for k, v in d.items()}) if isinstance(d, dict) else d


def synth_burst(n, H, W, seed=0, max_shift=2.0):
    from scipy.ndimage import gaussian_filter, shift as ndshift
    rng = np.random.default_rng(seed)
    up, pad = 4, 32
    scene = rng.random((3, H * up + 2 * pad, W * up + 2 * pad)).astype(np.float32)
    scene = np.stack([gaussian_filter(s, 3) for s in scene])
    scene = 0.05 + 0.8 * (scene - scene.min()) / (scene.max() - scene.min())
    frames, shifts = [], []
    for i in range(n):
        dy, dx = (0.0, 0.0) if i == 0 else rng.uniform(-max_shift, max_shift, 2)
        shifts.append((float(dy), float(dx)))
        sh = np.stack([ndshift(s, (dy * up, dx * up), order=1, mode="nearest") for s in scene])
        lr = sh[:, pad:pad + H * up, pad:pad + W * up].reshape(3, H, up, W, up).mean((2, 4))
        bay = np.empty((H, W), np.float32)
        bay[0::2, 0::2] = lr[0, 0::2, 0::2]
        bay[0::2, 1::2] = lr[1, 0::2, 1::2]
        bay[1::2, 0::2] = lr[1, 1::2, 0::2]
        bay[1::2, 1::2] = lr[2, 1::2, 1::2]
        bay = bay + np.sqrt(1.80710882e-4 * bay + 3.1937599182128e-6) * rng.standard_normal(bay.shape)
        frames.append(np.clip(bay, 0, 1).astype(np.float32))
    return np.stack(frames), shifts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--scale", type=float, default=1)
    ap.add_argument("--ts", type=int, default=32)
    ap.add_argument("--factors", type=str, default="1,2,2,2")
    ap.add_argument("--verbose", type=int, default=2)
    ap.add_argument("--out", type=str, default="")
    a = ap.parse_args()
    scale = int(a.scale) if a.scale == int(a.scale) else a.scale
    factors = [int(x) for x in a.factors.split(",")]

    install_shims()
    sys.path.insert(0, patched_reference_copy())
    import yaml
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution.params import update_snr_config, sanitize_config

    cfg = Cfg.wrap(yaml.safe_load(open(os.path.join(REFERENCE, "configs", "default.yaml"))))
    cfg.scale, cfg.verbose = scale, a.verbose
    bm = cfg.block_matching.tuning
    bm.tile_size, bm.factors = a.ts, factors
    L = len(factors)  # fine-to-coarse lists must all have len(factors) entries
    bm.tile_size_factors = [1] * (L - 1) + [0.5]
    bm.search_radii = [1] + [4] * (L - 1)
    bm.metrics = ["L1"] + ["L2"] * (L - 1)
    cfg.noise_model.alpha, cfg.noise_model.beta = 1.80710882e-4, 3.1937599182128e-6
    burst, shifts = synth_burst(a.n, a.size, a.size)
    std_curve = np.load(os.path.join(REFERENCE, "data", "noise_model_std_ISO_100.npy"))
    diff_curve = np.load(os.path.join(REFERENCE, "data", "noise_model_diff_ISO_100.npy"))
    brightness = float(np.mean(burst[0]))
    update_snr_config(cfg, brightness / std_curve[round(1000 * brightness)])  # as process() does
    cfg.exif = Cfg(cfa_pattern=[[0, 1], [1, 2]], iso=100, white_balance=[2.0, 1.0, 1.5, 0.0])
    cfg.noise_model.std_curve, cfg.noise_model.diff_curve = std_curve.tolist(), diff_curve.tolist()
    cfg.accumulated_robustness_denoiser.enabled = False
    sanitize_config(cfg, burst[0].shape)

    flows, robs = [], []
    _al, _cr = SR.align, SR.compute_robustness

    def cap_al(*x, **k):
        r = _al(*x, **k)
        flows.append(r.detach().clone().numpy())
        return r

    def cap_cr(*x, **k):
        r = _cr(*x, **k)
        robs.append(r.copy_to_host())
        return r
    SR.align, SR.compute_robustness = cap_al, cap_cr

    t0 = time.perf_counter()
    out, _ = SR.main(burst[0], burst[1:], cfg)
    dt = time.perf_counter() - t0
    out = out.copy_to_host()
    hr = out.shape[0] * out.shape[1]
    print("CUDASIM reference: %d frames %dx%d scale %s Ts %d -> %s in %.1f s = %.6f output MPix/s on %d host cores "
          "(GIL-bound)" % (a.n, a.size, a.size, scale, a.ts, out.shape, dt, hr / 1e6 / dt, os.cpu_count()))
    print("true (dy,dx):", shifts[1:], " median flow (dx,dy):",
          [np.median(f.reshape(-1, 2), 0).round(3).tolist() for f in flows])
    if a.out:
        np.savez(a.out, out=out, flows=np.stack(flows), robs=np.stack(robs), burst=burst, seconds=dt)


if __name__ == "__main__":
    main()
