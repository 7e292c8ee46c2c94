"""Golden-vector generator: runs the UNMODIFIED reference (Numba-CUDA + torch) on a real GPU and dumps every
stage's inputs/outputs as .npz fixtures.  Test tooling, not product code.

Where it runs: on the GPU box through gpurun, against `baseline/_ref/` (a verbatim, git-ignored copy of
/root/reference/{handheld_super_resolution,configs,data} that travels with the snapshot):

    cp -r /root/reference/{handheld_super_resolution,configs,data} baseline/_ref/      # once, in the container
    gpurun -- python tests/golden/make_golden_gpu.py                                    # writes gpurun_out/golden/
    cp gpurun_out/golden/*.npz tests/golden/                                            # commit the fixtures

Harness-level shims only (none alters what the reference computes):
  * stubs for modules the hot path never calls but imports (omegaconf, rawpy, exifread, imageio, skimage, matplotlib);
  * a bit-twiddling lowering of math.copysign(f64,f64): numba 0.65 + NVVM 12.9 reject libdevice's __nv_copysign
    ("Unsupported intrinsic: llvm.copysign.f64"), copysign is exact so results are unchanged.
"""
import importlib.util
import json
import math
import os
import sys
import time
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("HHSR_REF", os.path.join(ROOT, "baseline", "_ref"))
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)

for name in ["omegaconf", "rawpy", "exifread", "imageio", "skimage", "skimage.filters", "matplotlib",
             "matplotlib.pyplot"]:
    sys.modules[name] = mock.MagicMock()
sys.path.insert(0, REF)


def install_copysign_shim():
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


def load_synth():
    p = os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200", "handheld_super_resolution", "synthetic.py")
    spec = importlib.util.spec_from_file_location("hhsr_synthetic", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


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


def to_np(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy()
    if hasattr(x, "copy_to_host"):
        return x.copy_to_host()
    return np.asarray(x)


def make_config(scale, Ts, factors, ref_frame, std_curve, diff_curve, **over):
    import yaml
    from handheld_super_resolution.params import update_snr_config, sanitize_config
    c = Cfg.wrap(yaml.safe_load(open(os.path.join(REF, "configs", "default.yaml"))))
    c.scale, c.verbose = scale, 0
    bm = c.block_matching.tuning
    bm.tile_size, bm.factors = Ts, list(factors)
    L = len(factors)
    bm.tile_size_factors = [1] * (L - 1) + [0.5]
    bm.search_radii = [1] + [4] * (L - 1)
    bm.metrics = ["L1"] + ["L2"] * (L - 1)
    c.noise_model.alpha, c.noise_model.beta = 1.80710882e-4, 3.1937599182128e-6
    brightness = float(np.mean(ref_frame))
    update_snr_config(c, brightness / std_curve[round(1000 * brightness)])  # as process() does
    c.exif = Cfg(cfa_pattern=[[0, 1], [1, 2]], iso=100, white_balance=[2.0, 1.0, 1.5, 0.0])
    c.noise_model.std_curve, c.noise_model.diff_curve = std_curve.tolist(), diff_curve.tolist()
    c.accumulated_robustness_denoiser.enabled = False
    for k, v in over.items():
        node = c
        ks = k.split("__")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    sanitize_config(c, ref_frame.shape)
    return c


def cfg_summary(c):
    bm = c.block_matching.tuning
    mt = c.merging.tuning
    return dict(scale=c.scale, tile_size=bm.tile_size, tile_sizes=list(bm.tile_sizes), factors=list(bm.factors),
                search_radii=list(bm.search_radii), metrics=list(bm.metrics), n_iter=c.ica.tuning.n_iter,
                k_detail=mt.k_detail, k_denoise=mt.k_denoise, D_th=mt.D_th, D_tr=mt.D_tr,
                k_stretch=mt.k_stretch, k_shrink=mt.k_shrink, kernel=c.merging.kernel,
                selection_law=c.merging.selection_law, t=c.robustness.tuning.t, s1=c.robustness.tuning.s1,
                s2=c.robustness.tuning.s2, Mt=c.robustness.tuning.Mt, alpha=c.noise_model.alpha,
                beta=c.noise_model.beta)


def run_pipeline(burst, cfg, full):
    """Run reference main() with capture hooks.  `full`: keep every intermediate (small bursts only)."""
    from numba import cuda
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution import alignment as AL
    from handheld_super_resolution import robustness as RB
    cap = {}
    state = {"frame": 0, "grey": 0, "pyr": 0, "kern": 0}
    saved = {}

    def patch(mod, name, fn):
        saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, fn)

    o_grey, o_pyr, o_ia, o_al = SR.compute_grey_images, AL.build_gaussian_pyramid, SR.init_alignment, SR.align
    o_up, o_l2, o_l1, o_ica = AL.upscale_lvl, AL.align_lvl_block_matching_L2, AL.align_lvl_block_matching_L1, AL.align_lvl_ica
    o_ir, o_cr, o_ek, o_m, o_mr, o_div = SR.init_robustness, SR.compute_robustness, SR.estimate_kernels, SR.merge, SR.merge_ref, SR.divide
    o_thr = RB.robustness_threshold

    def grey(img, method):
        out = o_grey(img, method)
        cuda.synchronize()
        if method == "FFT" and (full or state["grey"] == 0):
            cap["grey_%d" % state["grey"]] = to_np(out)
        if method == "FFT":
            state["grey"] += 1
        return out

    def pyr(image, factors=[1, 2, 4, 4], kernel="gaussian"):
        out = o_pyr(image, factors, kernel)
        torch.cuda.synchronize()
        if full or state["pyr"] <= 1:
            for i, lvl in enumerate(out):  # coarse -> fine
                if full or i < len(out) - 1:
                    cap["pyr_%d_c%d" % (state["pyr"], i)] = to_np(lvl)
        state["pyr"] += 1
        return out

    def init_al(ref_grey, config):
        out = o_ia(ref_grey, config)
        cuda.synchronize()
        for i in range(len(out[0])):  # coarse -> fine
            if full or i < len(out[0]) - 1:
                cap["ref_gradx_c%d" % i] = to_np(out[3][i])
                cap["ref_grady_c%d" % i] = to_np(out[4][i])
            cap["ref_hessian_c%d" % i] = to_np(out[5][i])
        return out

    def align(*a, **k):
        state["frame"] += 1
        out = o_al(*a, **k)
        cuda.synchronize()
        cap["flow_f%d" % state["frame"]] = to_np(out)
        return out

    def up(alignments, npatchs, l, config):
        out = o_up(alignments, npatchs, l, config)
        cap["flow_f%d_l%d_up" % (state["frame"], l)] = to_np(out)
        return out

    def l2(tyled, fft, moving, alignment, l, config):
        cap["flow_f%d_l%d_in" % (state["frame"], l)] = to_np(alignment)
        o_l2(tyled, fft, moving, alignment, l, config)
        cuda.synchronize()
        cap["flow_f%d_l%d_bm" % (state["frame"], l)] = to_np(alignment)

    def l1(ref_lvl, moving, alignments, l, config):
        cap["flow_f%d_l%d_in" % (state["frame"], l)] = to_np(alignments)
        o_l1(ref_lvl, moving, alignments, l, config)
        cuda.synchronize()
        cap["flow_f%d_l%d_bm" % (state["frame"], l)] = to_np(alignments)

    def ica(ref_img, gx, gy, hess, moving, alignment, l, config):
        o_ica(ref_img, gx, gy, hess, moving, alignment, l, config)
        cuda.synchronize()
        cap["flow_f%d_l%d_ica" % (state["frame"], l)] = to_np(alignment)

    def init_rob(*a, **k):
        out = o_ir(*a, **k)
        cuda.synchronize()
        if full and out[0] is not None:
            cap["ref_means"], cap["ref_stds"] = to_np(out[0]), to_np(out[1])
        return out

    def thr(*a, **k):
        out = o_thr(*a, **k)
        cuda.synchronize()
        if full:
            cap["R_f%d" % state["frame"]] = to_np(out)
        return out

    def rob(*a, **k):
        out = o_cr(*a, **k)
        cuda.synchronize()
        cap["r_f%d" % state["frame"]] = to_np(out)
        return out

    def kern(img, config):
        out = o_ek(img, config)
        cuda.synchronize()
        state["kern"] += 1
        cap["covs_%d" % state["kern"]] = to_np(out)  # 1..N-1 comp frames in order, last = ref
        return out

    def merge(comp, al, covs, r, num, den, cfa, config):
        o_m(comp, al, covs, r, num, den, cfa, config)
        cuda.synchronize()
        if full and state["frame"] == len(burst) - 1:   # accumulators after the last comp frame
            cap["num_comp"], cap["den_comp"] = to_np(num), to_np(den)

    def merge_ref(ref, kernels, num, den, cfa, config, acc_rob=None):
        if acc_rob is not None:
            cap["acc_rob"] = to_np(acc_rob)
            o_mr(ref, kernels, num, den, cfa, config, acc_rob)
        else:
            o_mr(ref, kernels, num, den, cfa, config)
        cuda.synchronize()
        cap["num_final"], cap["den_final"] = to_np(num), to_np(den)

    for mod, name, fn in [(SR, "compute_grey_images", grey), (AL, "build_gaussian_pyramid", pyr),
                          (SR, "init_alignment", init_al), (SR, "align", align), (AL, "upscale_lvl", up),
                          (AL, "align_lvl_block_matching_L2", l2), (AL, "align_lvl_block_matching_L1", l1),
                          (AL, "align_lvl_ica", ica), (SR, "init_robustness", init_rob),
                          (RB, "robustness_threshold", thr), (SR, "compute_robustness", rob),
                          (SR, "estimate_kernels", kern), (SR, "merge", merge), (SR, "merge_ref", merge_ref)]:
        patch(mod, name, fn)
    try:
        out, _ = SR.main(burst[0], burst[1:], cfg)
        cuda.synchronize()
        if not full:
            cap["out"] = to_np(out)
        else:  # out == num_final / den_final exactly (one IEEE division, utils.py:84-90): store the check only
            with np.errstate(all="ignore"):
                o, q = to_np(out), cap["num_final"] / cap["den_final"]
            cap["out_is_num_over_den"] = np.array(bool(np.array_equal(o, q, equal_nan=True)))
    finally:
        for (mod, name), fn in saved.items():
            setattr(mod, name, fn)
    return cap


def crops(a, size=48):
    """Corner + centre crops of a [H,W,...] array (keeps big goldens small)."""
    H, W = a.shape[:2]
    cy, cx = (H - size) // 2, (W - size) // 2
    return dict(tl=a[:size, :size].copy(), tr=a[:size, W - size:].copy(), bl=a[H - size:, :size].copy(),
                br=a[H - size:, W - size:].copy(), c=a[cy:cy + size, cx:cx + size].copy())


def main():
    install_copysign_shim()
    synth = load_synth()
    from numba import cuda
    report = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "tf32_default": torch.backends.cudnn.allow_tf32}
    std_curve = np.load(os.path.join(REF, "data", "noise_model_std_ISO_100.npy"))
    diff_curve = np.load(os.path.join(REF, "data", "noise_model_diff_ISO_100.npy"))
    np.savez_compressed(os.path.join(OUT, "noise_curves_iso100.npz"), std_curve=std_curve, diff_curve=diff_curve)

    # ------------------------------------------------------------------ A. tiny burst, everything kept
    burst, shifts = synth.synth_burst(3, 118, 170, seed=1, max_shift=2.5, quantize_bits=14)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        cfg = make_config(2, 32, [1, 2, 2], burst[0], std_curve, diff_curve)
        t0 = time.perf_counter()
        cap = run_pipeline(burst, cfg, full=True)
        report["tiny_tf32_%s_s" % tf32] = time.perf_counter() - t0
        cap["burst"] = burst
        cap["shifts"] = np.array(shifts)
        cap["cfg_json"] = np.array(json.dumps(cfg_summary(cfg)))
        if not tf32:
            np.savez_compressed(os.path.join(OUT, "tiny_pipeline.npz"), **cap)
            tiny = cap
        else:
            report["tiny_tf32_vs_fp32"] = {k: float(np.nanmax(np.abs(cap[k].astype(np.float64) - tiny[k])))
                                           for k in cap if k.startswith(("flow_f", "pyr_", "out", "r_f"))
                                           and k in tiny and cap[k].shape == tiny[k].shape and cap[k].dtype.kind == "f"}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    # ------------------------------------------------------------------ B. medium burst, default factors [1,2,4,4]
    burst_m, shifts_m = synth.synth_burst(3, 700, 740, seed=2, max_shift=3.0, quantize_bits=14)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        cfg = make_config(2, 32, [1, 2, 4, 4], burst_m[0], std_curve, diff_curve)
        cap = run_pipeline(burst_m, cfg, full=False)
        keep = {k: v for k, v in cap.items() if k.startswith(("flow_", "ref_hessian"))}
        for k in ["out", "num_final", "den_final", "r_f1", "r_f2", "covs_1", "covs_3", "grey_0"]:
            for ck, cv in crops(cap[k]).items():
                keep["%s__%s" % (k, ck)] = cv
        keep["out_mean"] = np.nanmean(cap["out"].astype(np.float64), axis=(0, 1))
        keep["out_nan"] = np.array(int(np.isnan(cap["out"]).sum()))
        keep["r_mean"] = np.array([cap["r_f%d" % i].astype(np.float64).mean() for i in (1, 2)])
        if not tf32:
            keep["seed"] = np.array(2)
            keep["shifts"] = np.array(shifts_m)
            keep["cfg_json"] = np.array(json.dumps(cfg_summary(cfg)))
            keep["burst_u16"] = np.round(burst_m * 16383.0).astype(np.uint16)
            np.savez_compressed(os.path.join(OUT, "medium_pipeline.npz"), **keep)
            med = keep
        else:
            report["medium_tf32_vs_fp32"] = {k: float(np.nanmax(np.abs(keep[k].astype(np.float64) - med[k])))
                                             for k in keep if k.startswith(("flow_f", "out__"))}
    torch.backends.cudnn.allow_tf32 = False

    # ------------------------------------------------------------------ C. stage-level cases
    from handheld_super_resolution import merge as MG, kernels as KN, robustness as RB, ICA, block_matching as BM
    from handheld_super_resolution import alignment as AL
    st = {}
    raw = np.ascontiguousarray(burst[1][:48, :72])      # crop: 2x3 flow tiles of 32, keeps the fixture small
    ref0 = np.ascontiguousarray(burst[0][:48, :72])
    st["raw"], st["ref"] = raw, ref0
    d_raw = cuda.to_device(raw)
    d_ref = cuda.to_device(ref0)
    cfa = cuda.to_device(np.array([[0, 1], [1, 2]]))
    wb = cuda.to_device(np.array([2.0, 1.0, 1.5, 0.0]))
    H, W = raw.shape
    rng = np.random.default_rng(7)
    flow_irreg = (tiny["flow_f1"][:2, :3] + rng.uniform(-1.5, 1.5, (2, 3, 2))).astype(np.float32)
    flow_irreg[0, 0] = (-9.3, 7.1)
    flow_irreg[-1, -1] = (6.4, 11.7)
    st["flow_irreg"] = flow_irreg
    r_rand = rng.uniform(0, 1, raw.shape).astype(np.float32)
    st["r_rand"] = r_rand
    covs1 = np.ascontiguousarray(tiny["covs_1"][:24, :36])
    covs_ref = np.ascontiguousarray(tiny["covs_3"][:24, :36])
    st["covs1"], st["covs_ref"] = covs1, covs_ref
    for scale in (1, 1.5, 2, 3):
        for kern in ("steerable", "iso"):
            if kern == "iso" and scale not in (1.5, 2):
                continue
            cfg = make_config(scale, 32, [1, 2, 2], burst[0], std_curve, diff_curve, merging__kernel=kern)
            hs, ws = round(scale * H), round(scale * W)
            num = cuda.to_device(np.zeros((hs, ws, 3), np.float32))
            den = cuda.to_device(np.zeros((hs, ws, 3), np.float32))
            MG.merge(d_raw, cuda.to_device(flow_irreg), cuda.to_device(covs1), cuda.to_device(r_rand), num, den, cfa, cfg)
            cuda.synchronize()
            tag = "s%s_%s" % (str(scale).replace(".", "p"), kern)
            st["merge_num_" + tag], st["merge_den_" + tag] = to_np(num), to_np(den)
            MG.merge_ref(d_ref, cuda.to_device(covs_ref), num, den, cfa, cfg)
            cuda.synchronize()
            st["mergeref_num_" + tag], st["mergeref_den_" + tag] = to_np(num), to_np(den)
    # accumulated-robustness denoise mode of merge_ref (merge.py:54-65,167-176,223-233)
    acc_rob = np.round(rng.uniform(0, 4, raw.shape) * 4) / 4
    st["acc_rob"] = acc_rob
    cfg = make_config(2, 32, [1, 2, 2], burst[0], std_curve, diff_curve)
    cfg.accumulated_robustness_denoiser.enabled = True
    cfg.accumulated_robustness_denoiser.merge.enabled = True
    num = cuda.to_device(st["merge_num_s2_steerable"])
    den = cuda.to_device(st["merge_den_s2_steerable"])
    MG.merge_ref(d_ref, cuda.to_device(covs_ref), num, den, cfa, cfg, cuda.to_device(acc_rob))
    cuda.synchronize()
    st["mergeref_accrob_num"], st["mergeref_accrob_den"] = to_np(num), to_np(den)
    # kernels: hard threshold law, and a frame with flat + saturated patches (NaN covariances, Q5)
    raw_flat = raw.copy()
    raw_flat[4:20, 6:30] = 0.5
    raw_flat[24:44, 40:66] = 1.0
    raw_flat[30:46, 2:20] = 0.0
    st["raw_flat"] = raw_flat
    for law in ("linear", "hard_threshold"):
        cfg = make_config(2, 32, [1, 2, 2], burst[0], std_curve, diff_curve, merging__selection_law=law)
        st["covs_flat_" + law] = to_np(KN.estimate_kernels(cuda.to_device(raw_flat), cfg))
        st["covs_" + law] = to_np(KN.estimate_kernels(d_raw, cfg))
    # merge with NaN covariances
    cfg = make_config(2, 32, [1, 2, 2], burst[0], std_curve, diff_curve)
    num = cuda.to_device(np.zeros((2 * H, 2 * W, 3), np.float32))
    den = cuda.to_device(np.zeros((2 * H, 2 * W, 3), np.float32))
    MG.merge(cuda.to_device(raw_flat), cuda.to_device(flow_irreg), cuda.to_device(st["covs_flat_linear"]),
             cuda.to_device(r_rand), num, den, cfa, cfg)
    MG.merge_ref(cuda.to_device(raw_flat), cuda.to_device(st["covs_flat_linear"]), num, den, cfa, cfg)
    cuda.synchronize()
    st["merge_flat_num"], st["merge_flat_den"] = to_np(num), to_np(den)
    # robustness with irregular flow (exercises S, OOB warps)
    means, stds = RB.init_robustness(d_ref, cfa, wb, cfg)
    r_irreg = RB.compute_robustness(d_raw, means, stds, cuda.to_device(flow_irreg), cfa, wb,
                                    (cuda.to_device(std_curve), cuda.to_device(diff_curve)), cfg)
    st["r_irreg"] = to_np(r_irreg)
    guide = RB.compute_guide_image(d_raw, cfa, wb)
    lm, ls = RB.compute_local_stats(guide)
    st["guide_f1"], st["lmeans_f1"], st["lstds_f1"] = to_np(guide), to_np(lm), to_np(ls)
    st["S_irreg"] = to_np(RB.compute_s(cuda.to_device(flow_irreg), cfg.robustness.tuning.Mt,
                                       cfg.robustness.tuning.s1, cfg.robustness.tuning.s2))
    np.savez_compressed(os.path.join(OUT, "stage_cases.npz"), **st)

    # ------------------------------------------------------------------ D. alignment kernels per tile size
    al = {}
    g = torch.Generator(device="cuda").manual_seed(5)
    for ts in (8, 16, 32, 64):
        ny, nx = 5, 6
        h, w = ts * ny, ts * nx
        base = torch.rand((1, 1, h // 4 + 8, w // 4 + 8), device="cuda", generator=g)
        base = torch.nn.functional.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)[0, 0]
        ref = base[8:8 + h, 8:8 + w].contiguous()
        mov = (base[6:6 + h, 11:11 + w - 5] + 0.01 * torch.rand((h, w - 5), device="cuda", generator=g)).contiguous()
        flow0 = ((torch.rand((ny, nx, 2), device="cuda", generator=g) - 0.5) * 6).contiguous()
        al["ref_%d" % ts], al["mov_%d" % ts], al["flow0_%d" % ts] = to_np(ref), to_np(mov), to_np(flow0)
        cfg = make_config(2, 32, [1, 2, 2], burst[0], std_curve, diff_curve)
        cfg.block_matching.tuning.tile_sizes = [ts, ts, ts]
        gx, gy, hess = ICA.init_ica(ref, ts, cfg)
        cuda.synchronize()
        al["gx_%d" % ts], al["gy_%d" % ts], al["hess_%d" % ts] = to_np(gx), to_np(gy), to_np(hess)
        f = flow0.clone()
        ICA.align_lvl_ica(ref, gx, gy, hess, mov, f, 0, cfg)
        cuda.synchronize()
        al["ica_%d" % ts] = to_np(f)
        # L2 block matching, radius 4
        cfg.block_matching.tuning.search_radii = [4, 4, 4]
        tiled = ref.unfold(0, ts, ts).unfold(1, ts, ts)
        tiled = torch.nn.functional.pad(tiled, (4, 4, 4, 4), mode="constant", value=0)
        fft = torch.fft.rfft2(tiled, dim=(-2, -1))
        f = flow0.clone()
        BM.align_lvl_block_matching_L2(tiled, fft, mov, f, 0, cfg)
        torch.cuda.synchronize()
        al["bm2_%d" % ts] = to_np(f)
        if ts >= 16:
            cfg.block_matching.tuning.search_radii = [1, 1, 1]
            f = flow0.clone()
            BM.align_lvl_block_matching_L1(ref, mov, f, 0, cfg)
            cuda.synchronize()
            al["bm1_%d" % ts] = to_np(f)
    # flow upscaling modes
    cfg = make_config(2, 32, [1, 2, 4, 4], burst_m[0], std_curve, diff_curve)
    fl = torch.rand((5, 7, 2), device="cuda", generator=g) * 4 - 2
    al["up_in"] = to_np(fl)
    for mode in ("nearest", "bilinear", "bicubic"):
        cfg.block_matching.tuning.flow_upscale_mode = mode
        al["up_l2_" + mode] = to_np(AL.upscale_lvl(fl, (11, 15), 2, cfg))
        al["up_l0_" + mode] = to_np(AL.upscale_lvl(fl, (11, 15), 0, cfg))
    np.savez_compressed(os.path.join(OUT, "alignment_cases.npz"), **al)

    report["files"] = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
