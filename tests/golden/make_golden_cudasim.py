"""Golden-vector generator #2: runs the reference's own kernels under NUMBA_ENABLE_CUDASIM=1 in the build
container (no GPU) on tiny seeded inputs and stores inputs + outputs in tests/golden/cudasim_cases.npz.

    python tests/golden/make_golden_cudasim.py          # needs /root/reference; ~10 min on 8 cores

Test tooling, not product code.  It imports the reference from a patched scratch copy made by
baseline/run_reference_cudasim.py (4 Python-semantics patches + simulator shims documented there; none changes
what the compiled GPU path computes).  The simulator evaluates kernels with NumPy scalar semantics (NEP 50:
float32 (op) Python-float stays float32) whereas compiled Numba promotes to float64, so these vectors pin the
ALGORITHM (index conventions, borders, NaN/inf rules) to ~1e-5; the B200 goldens (make_golden_gpu.py) pin the
numerics.
"""
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import run_reference_cudasim as H  # noqa: E402


def main():
    H.install_shims()
    sys.path.insert(0, H.patched_reference_copy())
    from numba import cuda
    from handheld_super_resolution import merge as MG, kernels as KN, robustness as RB, ICA
    rng = np.random.default_rng(11)
    out = {}
    Hh, Ww, ts = 16, 24, 8
    sm = np.linspace(0, 1, Hh)[:, None] * 0.3 + np.linspace(0, 1, Ww)[None, :] * 0.4
    raw = (0.1 + sm + 0.15 * rng.random((Hh, Ww))).astype(np.float32)
    raw[2:6, 3:9] = 0.5          # flat patch -> NaN covariances with the linear law
    ref = (raw + 0.02 * rng.standard_normal((Hh, Ww))).astype(np.float32).clip(0, 1)
    flow = rng.uniform(-1.6, 1.6, (Hh // ts, Ww // ts, 2)).astype(np.float32)
    flow[0, 0] = (-3.2, 0.4)     # pushes part of tile 0 out of the frame
    r = rng.random((Hh, Ww)).astype(np.float32)
    cfa = np.array([[0, 1], [1, 2]])
    wb = np.array([2.0, 1.0, 1.5, 0.0])
    std_curve = np.load(os.path.join(H.REFERENCE, "data", "noise_model_std_ISO_100.npy"))
    diff_curve = np.load(os.path.join(H.REFERENCE, "data", "noise_model_diff_ISO_100.npy"))
    out.update(raw=raw, ref=ref, flow=flow, r=r, std_curve=std_curve, diff_curve=diff_curve)

    import yaml
    cfg = H.Cfg.wrap(yaml.safe_load(open(os.path.join(H.REFERENCE, "configs", "default.yaml"))))
    cfg.verbose = 0
    cfg.noise_model.alpha, cfg.noise_model.beta = 1.80710882e-4, 3.1937599182128e-6
    cfg.merging.tuning.k_detail, cfg.merging.tuning.k_denoise = 0.25, 3.0
    cfg.merging.tuning.D_th, cfg.merging.tuning.D_tr = 0.71, 1.0
    cfg.block_matching.tuning.tile_size = ts
    cfg.accumulated_robustness_denoiser.enabled = False

    # ---- kernel estimation (kernels.py:29-243)
    for law in ("linear", "hard_threshold"):
        cfg.merging.selection_law = law
        out["covs_" + law] = KN.estimate_kernels(cuda.to_device(raw), cfg).copy_to_host()
    cfg.merging.selection_law = "linear"
    covs = out["covs_linear"]
    covs_ref = KN.estimate_kernels(cuda.to_device(ref), cfg).copy_to_host()
    out["covs_ref"] = covs_ref

    # ---- merge (merge.py:236-434) and merge_ref (merge.py:22-233)
    for scale in (1, 1.5, 2):
        cfg.scale = scale
        hs, ws = round(scale * Hh), round(scale * Ww)
        num, den = cuda.to_device(np.zeros((hs, ws, 3), np.float32)), cuda.to_device(np.zeros((hs, ws, 3), np.float32))
        MG.merge(cuda.to_device(raw), cuda.to_device(flow), cuda.to_device(covs), cuda.to_device(r), num, den,
                 cuda.to_device(cfa), cfg)
        tag = str(scale).replace(".", "p")
        out["merge_num_s" + tag], out["merge_den_s" + tag] = num.copy_to_host(), den.copy_to_host()
        MG.merge_ref(cuda.to_device(ref), cuda.to_device(covs_ref), num, den, cuda.to_device(cfa), cfg)
        out["mergeref_num_s" + tag], out["mergeref_den_s" + tag] = num.copy_to_host(), den.copy_to_host()
        print("merge scale", scale, "done", flush=True)
    cfg.scale = 2

    # ---- robustness (robustness.py:23-170)
    means, stds = RB.init_robustness(cuda.to_device(ref), cuda.to_device(cfa), cuda.to_device(wb), cfg)
    out["ref_means"], out["ref_stds"] = means.copy_to_host(), stds.copy_to_host()
    rr = RB.compute_robustness(cuda.to_device(raw), means, stds, cuda.to_device(flow), cuda.to_device(cfa),
                               cuda.to_device(wb), (cuda.to_device(std_curve), cuda.to_device(diff_curve)), cfg)
    out["robustness"] = rr.copy_to_host()
    print("robustness done", flush=True)

    # ---- ICA (ICA.py:15-274), tile sizes 8 and 16, 2x2 tiles
    for t in (8, 16):
        h, w = 2 * t, 2 * t
        base = torch.rand((1, 1, h // 4 + 4, w // 4 + 4), generator=torch.Generator().manual_seed(t))
        base = torch.nn.functional.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)[0, 0]
        ref_i = base[4:4 + h, 4:4 + w].contiguous()
        mov_i = base[3:3 + h, 6:6 + w - 3].contiguous()
        fl0 = torch.tensor(rng.uniform(-2.5, 2.5, (2, 2, 2)).astype(np.float32))
        cfg.block_matching.tuning.tile_sizes = [t]
        gx, gy, hess = ICA.init_ica(ref_i, t, cfg)
        f = fl0.clone()
        ICA.align_lvl_ica(ref_i, gx, gy, hess, mov_i, f, 0, cfg)
        out["ica%d_ref" % t], out["ica%d_mov" % t], out["ica%d_flow0" % t] = ref_i.numpy(), mov_i.numpy(), fl0.numpy()
        out["ica%d_gx" % t], out["ica%d_gy" % t] = gx.copy_to_host(), gy.copy_to_host()
        out["ica%d_hess" % t], out["ica%d_flow" % t] = hess.copy_to_host(), f.numpy()
        print("ica", t, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "cudasim_cases.npz"), **out)
    print("wrote", os.path.join(HERE, "cudasim_cases.npz"))


if __name__ == "__main__":
    main()
