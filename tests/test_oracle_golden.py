"""CPU tests (`-m "not gpu"`): the NumPy oracle against the golden vectors — the pin that lets the oracle stand in
for the reference.  B200 goldens come from the unmodified reference run through Numba-CUDA
(tests/golden/make_golden_gpu.py); CUDASIM goldens from the reference kernels under the Numba simulator
(tests/golden/make_golden_cudasim.py)."""
import numpy as np
import pytest

import hhsr_oracle as O
from helpers import CFA, WB, curves, load, maxdiff, plain_cfg, reldiff


@pytest.fixture(scope="module")
def tiny():
    return load("tiny_pipeline.npz")


@pytest.fixture(scope="module")
def stage():
    return load("stage_cases.npz")


def test_grey_and_pyramid(tiny):
    cfg = plain_cfg(tiny["cfg_json"])
    for i in range(3):
        assert maxdiff(O.grey_fft(tiny["burst"][i]), tiny["grey_%d" % i]) < 2e-6
    ref = O.init_alignment(tiny["grey_0"], cfg)
    for i in range(3):
        assert maxdiff(ref["pyramid"][i], tiny["pyr_0_c%d" % i]) < 1e-6
        assert maxdiff(ref["gradx"][i], tiny["ref_gradx_c%d" % i]) < 1e-6
        assert maxdiff(ref["grady"][i], tiny["ref_grady_c%d" % i]) < 1e-6
        assert reldiff(ref["hessian"][i], tiny["ref_hessian_c%d" % i], floor=1.0) < 5e-6


def test_grey_band_weights_equal_full_mask():
    """The half-spectrum symmetrised mask used by the CUDA path reproduces the reference's full-spectrum mask."""
    rng = np.random.default_rng(0)
    for shape in [(48, 64), (50, 66), (37, 51), (46, 62)]:
        img = rng.random(shape).astype(np.float32)
        H, W = shape
        ky, kx = np.arange(H), np.arange(W // 2 + 1)
        my, myn = O.grey_band_weights(H)
        mx, mxn = O.grey_band_weights(W)
        m = 0.5 * (my[:, None] & mx[None, kx]) + 0.5 * (myn[:, None] & mxn[None, kx])
        g = np.fft.irfft2(np.fft.rfft2(img.astype(np.float64)) * m, s=shape)
        assert np.abs(g - O.grey_fft(img)).max() < 1e-6


def test_alignment_chain(tiny):
    cfg = plain_cfg(tiny["cfg_json"])
    ref = O.init_alignment(tiny["grey_0"], cfg)
    for f in (1, 2):
        tr = {}
        flow = O.align(ref, tiny["grey_%d" % f], cfg, tr)
        for l in (2, 1, 0):
            assert np.array_equal(np.rint(tr["l%d_bm" % l] - tr["l%d_in" % l]),
                                  np.rint(tiny["flow_f%d_l%d_bm" % (f, l)] - tiny["flow_f%d_l%d_in" % (f, l)])), \
                "integer block-matching offsets differ at level %d" % l
            assert maxdiff(tr["l%d_ica" % l], tiny["flow_f%d_l%d_ica" % (f, l)]) < 1e-5
        assert maxdiff(flow, tiny["flow_f%d" % f]) < 1e-5


@pytest.mark.parametrize("ts", [8, 16, 32, 64])
def test_alignment_kernels_per_tile_size(ts):
    c = load("alignment_cases.npz")
    ref, mov, flow0 = c["ref_%d" % ts], c["mov_%d" % ts], c["flow0_%d" % ts]
    gx, gy, hess = O.init_ica(ref, ts)
    assert maxdiff(gx, c["gx_%d" % ts]) < 1e-6 and maxdiff(gy, c["gy_%d" % ts]) < 1e-6
    assert reldiff(hess, c["hess_%d" % ts], floor=1.0) < 5e-6
    out = O.ica(ref, c["gx_%d" % ts], c["gy_%d" % ts], c["hess_%d" % ts], mov, flow0, ts, 3)
    assert maxdiff(out, c["ica_%d" % ts]) < 2e-5
    bm, margin = O.bm_l2(ref, mov, flow0, ts, 4, return_margin=True)
    same = np.all(bm == c["bm2_%d" % ts], axis=-1)
    assert np.all(same | (margin < 1e-6)), "L2 offsets differ on tiles with a clear minimum"
    if ts in (32, 64):
        assert np.array_equal(O.bm_l1_compiled(flow0), c["bm1_%d" % ts])     # SURVEY Q1 on real hardware
    if ts == 16:
        # documents SURVEY Q2: the reference's ts=16 L1 kernel is NOT rint (undefined behaviour upstream)
        assert not np.array_equal(O.bm_l1_compiled(flow0), c["bm1_16"])


def test_upscale_lvl_modes():
    c = load("alignment_cases.npz")
    for mode in ("nearest", "bilinear", "bicubic"):
        for l in (2, 0):
            up = O.upscale_lvl(c["up_in"], (11, 15), l, [32, 32, 32, 16], [1, 2, 4, 4], mode)
            assert maxdiff(up, c["up_l%d_%s" % (l, mode)]) < 1e-5


def test_kernels(tiny, stage):
    cfg = plain_cfg(tiny["cfg_json"])
    for k, img in ((1, tiny["burst"][1]), (2, tiny["burst"][2]), (3, tiny["burst"][0])):
        c = O.estimate_kernels(img, cfg)
        assert maxdiff(c, tiny["covs_%d" % k]) < 1e-6 and reldiff(c, tiny["covs_%d" % k]) < 5e-6
    for law in ("linear", "hard_threshold"):
        cfg = plain_cfg(tiny["cfg_json"], selection_law=law)
        assert maxdiff(O.estimate_kernels(stage["raw_flat"], cfg), stage["covs_flat_" + law]) < 1e-6
        assert maxdiff(O.estimate_kernels(stage["raw"], cfg), stage["covs_" + law]) < 1e-6
    assert np.isnan(stage["covs_flat_linear"]).any() and not np.isnan(stage["covs_flat_hard_threshold"]).any()


def test_robustness(tiny, stage):
    cfg = plain_cfg(tiny["cfg_json"])
    std, diff = curves()
    m, s = O.init_robustness(tiny["burst"][0], CFA, WB)
    assert maxdiff(m, tiny["ref_means"]) < 1e-7 and maxdiff(s, tiny["ref_stds"]) < 1e-7
    for f in (1, 2):
        r, R = O.compute_robustness(tiny["burst"][f], tiny["ref_means"], tiny["ref_stds"], tiny["flow_f%d" % f],
                                    CFA, WB, std, diff, cfg, return_R=True)
        assert maxdiff(R, tiny["R_f%d" % f]) < 2e-6 and maxdiff(r, tiny["r_f%d" % f]) < 2e-6
    m, s = O.init_robustness(stage["ref"], CFA, WB)
    r = O.compute_robustness(stage["raw"], m, s, stage["flow_irreg"], CFA, WB, std, diff, cfg)
    assert maxdiff(r, stage["r_irreg"]) < 2e-6
    assert maxdiff(O.compute_s(stage["flow_irreg"], 0.8, 2, 12), stage["S_irreg"]) == 0
    g = O.guide_image(stage["raw"], CFA, WB)
    assert maxdiff(g, stage["guide_f1"]) == 0
    lm, ls = O.local_stats(g)
    assert maxdiff(lm, stage["lmeans_f1"]) < 1e-7 and maxdiff(ls, stage["lstds_f1"]) < 1e-7


@pytest.mark.parametrize("scale,kern", [(1, "steerable"), (1.5, "steerable"), (2, "steerable"), (3, "steerable"),
                                        (1.5, "iso"), (2, "iso")])
def test_merge(stage, scale, kern):
    H, W = stage["raw"].shape
    num = np.zeros((round(scale * H), round(scale * W), 3), np.float32)
    den = np.zeros_like(num)
    O.accumulate(stage["raw"], stage["flow_irreg"], stage["covs1"], stage["r_rand"], num, den, CFA, scale, 32, kern == "iso")
    tag = "s%s_%s" % (str(scale).replace(".", "p"), kern)
    tol = 5e-6   # accumulators reach ~5 (float32 ulp 4.8e-7); the reference rounds to float32 after every tap
    assert maxdiff(num, stage["merge_num_" + tag]) < tol and maxdiff(den, stage["merge_den_" + tag]) < tol
    O.accumulate_ref(stage["ref"], stage["covs_ref"], num, den, CFA, scale, kern == "iso")
    assert maxdiff(num, stage["mergeref_num_" + tag]) < tol and maxdiff(den, stage["mergeref_den_" + tag]) < tol


def test_merge_special_modes(stage):
    num, den = stage["merge_num_s2_steerable"].copy(), stage["merge_den_s2_steerable"].copy()
    O.accumulate_ref(stage["ref"], stage["covs_ref"], num, den, CFA, 2, False, stage["acc_rob"], 2, 2, 8)
    assert maxdiff(num, stage["mergeref_accrob_num"]) < 2e-5 and maxdiff(den, stage["mergeref_accrob_den"]) < 2e-5
    H, W = stage["raw_flat"].shape
    num = np.zeros((2 * H, 2 * W, 3), np.float32)
    den = np.zeros_like(num)
    O.accumulate(stage["raw_flat"], stage["flow_irreg"], stage["covs_flat_linear"], stage["r_rand"], num, den, CFA, 2, 32)
    O.accumulate_ref(stage["raw_flat"], stage["covs_flat_linear"], num, den, CFA, 2)
    assert maxdiff(num, stage["merge_flat_num"]) < 5e-6 and maxdiff(den, stage["merge_flat_den"]) < 5e-6


def test_main_tiny(tiny):
    out, dbg = O.main(tiny["burst"][0], tiny["burst"][1:], plain_cfg(tiny["cfg_json"]))
    with np.errstate(all="ignore"):
        want = tiny["num_final"] / tiny["den_final"]
    assert bool(tiny["out_is_num_over_den"])
    assert maxdiff(out, want) < 2e-5
    assert maxdiff(dbg["num"], tiny["num_final"]) < 3e-5 and maxdiff(dbg["den"], tiny["den_final"]) < 3e-5
    assert maxdiff(dbg["accumulated robustness"], tiny["acc_rob"]) < 1e-5
    assert np.isnan(want).sum() > 0                                           # SURVEY Q7 is real


def test_cudasim_cases():
    """Same stages against the reference executed by the Numba simulator (float32 (op) Python float stays float32
    there, so tolerances are looser than against the B200 goldens)."""
    c = load("cudasim_cases.npz")
    cfg = plain_cfg(tile_size=8, selection_law="linear")
    for law in ("linear", "hard_threshold"):
        cfg["merging"]["selection_law"] = law
        cv = O.estimate_kernels(c["raw"], cfg)
        assert reldiff(cv, c["covs_" + law], floor=1e-2) < 2e-3
        assert np.array_equal(np.isnan(cv), np.isnan(c["covs_" + law]))
    for scale in (1, 1.5, 2):
        H, W = c["raw"].shape
        num = np.zeros((round(scale * H), round(scale * W), 3), np.float32)
        den = np.zeros_like(num)
        O.accumulate(c["raw"], c["flow"], c["covs_linear"], c["r"], num, den, CFA, scale, 8)
        tag = str(scale).replace(".", "p")
        assert maxdiff(num, c["merge_num_s" + tag]) < 1e-4 and maxdiff(den, c["merge_den_s" + tag]) < 1e-4
        O.accumulate_ref(c["ref"], c["covs_ref"], num, den, CFA, scale)
        assert maxdiff(num, c["mergeref_num_s" + tag]) < 1e-4 and maxdiff(den, c["mergeref_den_s" + tag]) < 1e-4
    m, s = O.init_robustness(c["ref"], CFA, WB)
    assert maxdiff(m, c["ref_means"]) < 1e-6 and maxdiff(s, c["ref_stds"]) < 1e-6
    r = O.compute_robustness(c["raw"], c["ref_means"], c["ref_stds"], c["flow"], CFA, WB, c["std_curve"], c["diff_curve"], cfg)
    assert maxdiff(r, c["robustness"]) < 1e-4
    for t in (8, 16):
        gx, gy, hess = O.init_ica(c["ica%d_ref" % t], t)
        assert maxdiff(gx, c["ica%d_gx" % t]) < 1e-6 and reldiff(hess, c["ica%d_hess" % t], floor=1.0) < 1e-5
        out = O.ica(c["ica%d_ref" % t], gx, gy, c["ica%d_hess" % t], c["ica%d_mov" % t], c["ica%d_flow0" % t], t, 3)
        assert maxdiff(out, c["ica%d_flow" % t]) < 1e-4


def config1_cfg(maker, burst):
    """BASELINE config 1 exactly as baseline/run_reference_cudasim.py builds it: scale 1, Ts 32, factors [1,2,2,2],
    metrics [L1,L2,L2,L2], radii [1,4,4,4], SNR-derived merge constants (process() semantics)."""
    from handheld_super_resolution.config import load_config
    from handheld_super_resolution.params import update_snr_config
    std, _ = curves()
    b = float(np.mean(burst[0]))
    full = load_config(overrides={"scale": 1, "block_matching": {"tuning": {"tile_size": 32}}})
    update_snr_config(full, b / std[round(1000 * b)])
    mt = full.merging.tuning
    return maker(scale=1, tile_size=32, tile_sizes=[32, 32, 32, 16], factors=[1, 2, 2, 2], search_radii=[1, 4, 4, 4],
                 metrics=["L1", "L2", "L2", "L2"], k_detail=mt.k_detail, k_denoise=mt.k_denoise, D_th=mt.D_th, D_tr=mt.D_tr)


def check_config1(out, flow, r, z):
    """Whole-pipeline comparison with the reference's OWN main() run under the Numba simulator (BASELINE config 1).  The
    simulator evaluates kernels with NumPy scalar semantics (float32 op Python-float stays float32) where compiled Numba
    promotes to float64, hence a handful of pixels (8 of 196 590 for the oracle) differ by up to ~1.4e-3."""
    assert np.abs(flow - z["flows"][0]).max() < 1e-5
    assert np.abs(r - z["robs"][0]).max() < 5e-5
    assert np.array_equal(np.isnan(out), np.isnan(z["out"]))
    m = np.isfinite(out)
    d = np.abs(out[m].astype(np.float64) - z["out"][m])
    assert np.percentile(d, 99.9) < 5e-5 and d.max() < 5e-3 and (d > 1e-4).sum() <= 40
    return float(d.max()), int((d > 1e-4).sum())


def test_oracle_against_reference_main_under_cudasim_config1():
    z = load("config1_cudasim.npz")
    out, dbg = O.main(z["burst"][0], z["burst"][1:], config1_cfg(plain_cfg, z["burst"]))
    check_config1(out, dbg["flow"][0], dbg["robustness"][0], z)
