"""Output side: raw2rgb.postprocess, the 8/16-bit quantisation and the frame-count
denoisers.  CPU part: the oracle against goldens made from the reference's own code (tests/golden/make_golden_post.py).
GPU part (`-m gpu`): the CUDA kernels, through the product's Python surface, against the same goldens and the oracle."""
import numpy as np
import pytest
import torch

import hhsr_oracle as O
from helpers import load

SHARP_DEFAULT = dict(enabled=True, amount=1.5, radius=3)
CASES = {   # name -> (do_color_correction, do_gamma, sharpening, do_devignette, needs xyz2cam)
    "post_default": (False, True, SHARP_DEFAULT, False),
    "post_ccm_devignette": (True, True, dict(enabled=True, amount=0.8, radius=2), True),
    "post_plain": (False, False, dict(enabled=False), False),
    "post_gamma_only": (False, True, None, False),
}


def nandiff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(a)
    return float(np.abs(a[m] - b[m]).max())


@pytest.fixture(scope="module")
def post():
    return load("post_cases.npz")


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_postprocess_against_reference(post, name):
    ccm, gamma, sharp, dev = CASES[name]
    got = O.postprocess(post["img"], ccm, False, gamma, sharp, dev, post["xyz2cam"])
    assert nandiff(got, post[name]) < (5e-6 if ccm else 1e-7)     # float32 matmul order of the colour matrix
    if name == "post_default":
        assert np.isnan(post[name]).sum() > 100                  # the NaN pixels spread through the 25-tap blur


def test_oracle_gaussian_equals_scipy():
    from scipy import ndimage as ndi
    rng = np.random.default_rng(3)
    for shape, sigma in (((40, 57), 3), ((7, 9), 3), ((33, 20), 1.5)):
        img = rng.random(shape).astype(np.float32)
        assert np.abs(O.gaussian_filter_reflect(img, sigma) - ndi.gaussian_filter(img, sigma, mode="reflect", truncate=4.0)).max() < 1e-7


def test_oracle_median_against_reference_kernel(post):
    s, rmax, mfc = post["median_params"]
    got = O.frame_count_denoising_median(post["median_noisy"], post["median_r_acc"], s, rmax, mfc)
    assert np.array_equal(got, post["median_out"])
    assert (post["median_out"] != post["median_noisy"]).mean() > 0.3      # the case exercises real windows


def test_host_helpers_match_oracle(post):
    from handheld_super_resolution import raw2rgb
    from handheld_super_resolution.utils_image import apply_orientation
    assert np.allclose(np.linalg.inv(raw2rgb.get_color_matrix(None, post["xyz2cam"])), O.color_matrix(post["xyz2cam"]), rtol=0, atol=0)
    r, w = raw2rgb.gaussian_taps(3)
    assert r == 12 and np.array_equal(w, O.gaussian_taps(3, 12))
    a = np.arange(24).reshape(2, 4, 3)
    assert np.array_equal(apply_orientation(a, 1), a)
    assert np.array_equal(apply_orientation(a, 3), a[::-1, ::-1])
    assert apply_orientation(a, 6).shape == (4, 2, 3) and np.array_equal(apply_orientation(a, 6)[0, 0], a[1, 0])
    assert np.array_equal(apply_orientation(a, 8)[0, 0], a[0, 3])


# ------------------------------------------------------------------------------------------------ GPU
def _sharp(d):
    from handheld_super_resolution.config import Config
    return None if d is None else Config.wrap(dict(d))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_postprocess_against_reference(post, name):
    from handheld_super_resolution import raw2rgb
    ccm, gamma, sharp, dev = CASES[name]
    img = torch.from_numpy(post["img"]).cuda()
    got = raw2rgb.postprocess(None, img.clone(), ccm, False, gamma, _sharp(sharp), dev, post["xyz2cam"]).cpu().numpy()
    # float64 blur like scipy; powf (<= 4 ulp) and, after devignetting, float32 where the reference keeps float64
    assert nandiff(got, post[name]) < (5e-6 if ccm else 1e-6)
    # quantised outputs: what run_handheld.py saves (nan_to_num, clip, rint(x * 255)); a value sitting on a rounding
    # boundary may fall to the other side (powf)
    for dtype, top in (("uint8", 255), ("uint16", 65535)):
        q = raw2rgb.postprocess(None, img.clone(), ccm, False, gamma, _sharp(sharp), dev, post["xyz2cam"], output_dtype=dtype)
        assert str(q.dtype) == "torch." + dtype
        want = O.img_as_ubyte(post[name], top).astype(np.int64)
        d = np.abs(q.cpu().numpy().astype(np.int64) - want)
        assert d.max() <= 1 and (d > 0).mean() < (2e-2 if top == 65535 else 1e-3)


@pytest.mark.gpu
def test_postprocess_large_against_oracle():
    """A 300 x 420 image (several CTAs per axis, width not a multiple of 4 floats per thread group)."""
    from handheld_super_resolution import raw2rgb
    rng = np.random.default_rng(5)
    img = rng.random((300, 421, 3)).astype(np.float32)
    img[299, 100:104, 1] = np.nan
    want = O.postprocess(img, False, False, True, SHARP_DEFAULT, False, None)
    got = raw2rgb.postprocess(None, torch.from_numpy(img).cuda(), False, False, True, _sharp(SHARP_DEFAULT), False, None)
    assert nandiff(got.cpu().numpy(), want) < 1e-6


@pytest.mark.gpu
def test_frame_count_denoisers(post):
    from handheld_super_resolution.config import Config
    from handheld_super_resolution.utils_image import frame_count_denoising_gauss, frame_count_denoising_median
    s, rmax, mfc = post["median_params"]
    noisy, r_acc = torch.from_numpy(post["median_noisy"]).cuda(), torch.from_numpy(post["median_r_acc"]).cuda()
    got = frame_count_denoising_median(noisy, r_acc, Config.wrap({"radius_max": rmax, "max_frame_count": mfc}), scale=s)
    assert np.array_equal(got.cpu().numpy(), post["median_out"])            # the reference's own kernel (CUDASIM)
    got = frame_count_denoising_gauss(noisy, r_acc, Config.wrap({"sigma_max": 1.5, "max_frame_count": mfc}), scale=s)
    want = O.frame_count_denoising_gauss(post["median_noisy"], post["median_r_acc"], s, 1.5, mfc)
    assert np.abs(got.cpu().numpy() - want).max() < 1e-6
    assert (want != post["median_noisy"]).mean() > 0.3
    with pytest.raises(RuntimeError):                                         # radius above the reference's window buffer
        frame_count_denoising_median(noisy, r_acc, Config.wrap({"radius_max": 9, "max_frame_count": mfc}), scale=s)


@pytest.mark.gpu
def test_process_with_postprocessing_and_denoisers(tmp_path):
    """process() end to end with the reference's default post-processing (unsharp mask + gamma), a frame-count
    denoiser and 8-bit output: equals the stages applied by hand to main()'s image; EXIF orientation honoured."""
    from handheld_super_resolution import main, process, raw2rgb
    from handheld_super_resolution.config import Config, load_config
    from handheld_super_resolution.synthetic import ALPHA_ISO100, BETA_ISO100, synth_burst
    from handheld_super_resolution.utils_image import frame_count_denoising_gauss
    from helpers import curves
    burst, _ = synth_burst(3, 96, 128, seed=9, max_shift=2.0, quantize_bits=12)
    std, diff = curves()
    np.savez(tmp_path / "burst.npz", burst=burst, cfa_pattern=[[0, 1], [1, 2]], white_balance=[2.0, 1.0, 1.5, 0.0],
             alpha=ALPHA_ISO100, beta=BETA_ISO100, std_curve=std, diff_curve=diff, orientation=6)

    def cfg():
        c = load_config(overrides={"scale": 2, "verbose": 0})
        bm = c.block_matching.tuning
        bm.tile_size, bm.factors, bm.tile_size_factors = 16, [1, 2, 2], [1, 1, 0.5]
        bm.search_radii, bm.metrics = [2, 4, 4], ["L2", "L2", "L2"]
        c.accumulated_robustness_denoiser.gauss.enabled = True
        return c
    c1 = cfg()
    assert c1.postprocessing.enabled and c1.postprocessing.sharpening.enabled      # the reference's defaults
    img8, dbg = process(str(tmp_path), c1, output_dtype="uint8")
    assert img8.dtype == np.uint8 and img8.shape == (256, 192, 3)                    # orientation 6: rotated 90 degrees CW
    imgf, _ = process(str(tmp_path), cfg())
    assert imgf.dtype == np.float32 and imgf.shape == (256, 192, 3)
    # by hand: c1 now carries the SNR-derived parameters process() wrote into it
    out, d2 = main(burst[0], burst[1:], c1)
    out = frame_count_denoising_gauss(out, d2["accumulated robustness"], c1.accumulated_robustness_denoiser.gauss, scale=2)
    p = c1.postprocessing
    want = raw2rgb.postprocess(None, out, p.do_color_correction, p.do_tonemapping, p.do_gamma_correction, p.sharpening,
                               p.do_devignetting, np.zeros((3, 3))).cpu().numpy()
    assert np.array_equal(np.rot90(want, k=-1, axes=(0, 1)), imgf, equal_nan=True)
    assert np.abs(img8.astype(np.int64) - O.img_as_ubyte(imgf).astype(np.int64)).max() == 0
    assert dbg["accumulated robustness"].shape == (128, 96)
