"""Golden vectors for the OUTPUT side, made in the build container (CPU only):

  * raw2rgb.postprocess — the reference's OWN function (/root/reference/handheld_super_resolution/raw2rgb.py:212-250)
    executed on a small linear RGB image with NaN pixels, for the default configuration (unsharp mask amount 1.5 radius 3,
    gamma) and for colour correction + devignetting.  scikit-image is not installed here, so the one call the reference
    makes into it — skimage.filters.unsharp_mask(img, radius, amount, channel_axis=2, preserve_range=True) — is restated
    (scikit-image 0.2x, filters/_unsharp_mask.py) on top of the REAL scipy.ndimage.gaussian_filter it calls;
  * frame_count_denoising_median — the reference's Numba kernel (utils_image.py:251-315) run under NUMBA_ENABLE_CUDASIM,
    launched directly (its host wrapper reads `config.mode` / `config.scale` from the denoiser sub-config, which does not
    have them).  The Gaussian variant cannot execute at all upstream (range() over a float) and has no golden.

    python tests/golden/make_golden_post.py      # writes tests/golden/post_cases.npz (needs /root/reference)

Test tooling, not product code.
"""
import os
import sys
from unittest import mock

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402
from scipy import ndimage as ndi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import run_reference_cudasim as H  # noqa: E402


def unsharp_mask(image, radius=1.0, amount=1.0, preserve_range=False, channel_axis=None):
    assert preserve_range and channel_axis == 2 and image.dtype == np.float32
    out = np.empty_like(image)
    for c in range(image.shape[2]):
        blurred = ndi.gaussian_filter(image[..., c], sigma=radius, mode="reflect", truncate=4.0)
        out[..., c] = image[..., c] + (image[..., c] - blurred) * amount
    return out


class Sharp(dict):
    __getattr__ = dict.__getitem__


def main():
    H.install_shims()
    sys.path.insert(0, H.patched_reference_copy())
    for name in ["skimage.filters"]:
        sys.modules[name] = mock.MagicMock()
    from handheld_super_resolution import raw2rgb as R
    from handheld_super_resolution import utils_image as UI
    R.filters.unsharp_mask = unsharp_mask
    R.img_as_float32 = lambda x: np.asarray(x, np.float32)
    rng = np.random.default_rng(21)
    out = {}
    h, w = 44, 61
    base = rng.random((h // 4 + 2, w // 4 + 2, 3))
    img = np.kron(base, np.ones((4, 4, 1)))[:h, :w] * 0.7 + 0.3 * rng.random((h, w, 3))
    img = (img * 1.1 - 0.03).astype(np.float32)           # a few values outside [0, 1]
    img[h - 1, 5:9, 0] = np.nan                           # NaN pixels as main() leaves them on the last row / column
    img[7:11, w - 1, 2] = np.nan
    out["img"] = img
    xyz2cam = np.array([[1.0234, -0.2969, -0.2266], [-0.5625, 1.6328, -0.0469], [-0.0703, 0.2188, 0.6406]], np.float32)
    out["xyz2cam"] = xyz2cam
    # default configuration of configs/default.yaml:44-53: no colour correction, unsharp mask, gamma
    out["post_default"] = R.postprocess(object(), img.copy(), False, False, True, Sharp(enabled=True, amount=1.5, radius=3), False, None)
    out["post_ccm_devignette"] = R.postprocess(object(), img.copy(), True, False, True, Sharp(enabled=True, amount=0.8, radius=2), True, xyz2cam)
    out["post_plain"] = R.postprocess(object(), img.copy(), False, False, False, Sharp(enabled=False), False, None)
    out["post_gamma_only"] = R.postprocess(object(), img.copy(), False, False, True, None, False, None)

    # frame-count median denoiser, kernel launched directly (threads per block as the wrapper uses them, utils_image.py:247-250)
    from numba import cuda
    hs, ws, scale = 20, 28, 2
    noisy = rng.random((hs, ws, 3)).astype(np.float32)
    r_acc = np.round(rng.uniform(0, 10, (hs // scale, ws // scale)) * 2) / 2       # [H, W] float64, some above max_frame_count
    den = cuda.device_array(noisy.shape, np.float32)
    UI.cuda_frame_count_denoising_median[(2, 2, 3), (16, 16, 1)](cuda.to_device(noisy), den, cuda.to_device(r_acc), scale, 3, 8, False)
    out["median_noisy"], out["median_r_acc"], out["median_out"] = noisy, r_acc, den.copy_to_host()
    out["median_params"] = np.array([scale, 3, 8], np.float64)
    np.savez_compressed(os.path.join(HERE, "post_cases.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
