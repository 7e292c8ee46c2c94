"""Host-side parameter derivation — mirrors handheld_super_resolution/params.py of the reference
(update_snr_config :59-96, sanitize_config :4-57, lerp :99-123): same names, same mutations of `config`, same
exceptions.  One deliberate addition: sanitize_config also checks the true pyramid geometry (valid convolutions
shrink each level by 2*radius before subsampling), which the reference's check under-estimates and
then fails later with an empty tile grid."""
import numpy as np


def lerp(x, x_range, y_range):
    """params.py:99-123."""
    x0, x1 = x_range
    y0, y1 = y_range
    assert x0 < x1
    assert y0 != y1
    t = max(0.0, min(1.0, (x - x0) / (x1 - x0)))
    return y0 + (y1 - y0) * t


def update_snr_config(config, SNR):
    """params.py:59-96: SNR in [6,30] -> tile size {64,32,16} and the merge tuning constants."""
    SNR = float(np.clip(SNR, 6, 30))
    Ts = 64 if SNR <= 14 else (32 if SNR <= 22 else 16)
    bm = config.block_matching.tuning
    if bm.tile_size != "SNR_based":
        assert isinstance(bm.tile_size, int), "tile_size should be an integer or 'SNR_based'"
        Ts = bm.tile_size
    else:
        bm.tile_size = Ts
    bm.tile_sizes = [int(Ts * s) for s in bm.tile_size_factors]
    mt = config.merging.tuning
    for key, rng in (("k_detail", [0.33, 0.25]), ("k_denoise", [5.0, 3.0]), ("D_th", [0.81, 0.71]),
                     ("D_tr", [1.24, 1])):
        if mt[key] == "SNR_based":
            mt[key] = lerp(SNR, [6, 30], rng)
        else:
            assert isinstance(mt[key], float), "%s should be a float or 'SNR_based'" % key


def pyramid_shapes(shape, factors):
    """Level shapes fine -> coarse of build_gaussian_pyramid (utils_image.py:360-391): valid Gaussian of
    radius int(2f+0.5), then stride f."""
    h, w = shape
    out = []
    for f in factors:
        if f != 1:
            r = int(4 * f * 0.5 + 0.5)
            h, w = (h - 2 * r) // f, (w - 2 * r) // f
        out.append((h, w))
    return out


def sanitize_config(config, imshape):
    """params.py:4-57."""
    if config.mode == "grey" and config.grey_method != "FFT":
        raise NotImplementedError("Grey level images should be obtained with FFT")
    assert config.scale >= 1
    ard = config.accumulated_robustness_denoiser
    if not config.robustness.enabled and (ard.median.enabled or ard.gauss.enabled or ard.merge.enabled):
        raise ValueError("Accumulated robustness denoiser cannot be enabled if robustness is disabled.")
    if not config.robustness.enabled and config.robustness.save_mask:
        raise ValueError("Robustness mask cannot be saved if robustness is disabled.")
    assert config.merging.kernel in ["steerable", "iso"], f"Unknown kernel type {config.merging.kernel}"
    assert config.mode in ["bayer", "grey"], f"Unknown mode {config.mode}"
    if sum(1 if x.enabled else 0 for x in (ard.median, ard.gauss, ard.merge)) > 1:
        raise ValueError("Only one accumulated robustness denoiser can be enabled at a time.")
    assert config.ica.tuning.n_iter > 0, "Number of ICA iterations should be positive."
    assert config.ica.tuning.sigma_blur >= 0, f"Invalid sigma blur {config.ica.tuning.sigma_blur}."
    assert len(imshape) == 2, f"Input image shape should be 2D, got {imshape}."
    bm = config.block_matching.tuning
    Ts = bm.tile_size
    padded = (Ts * int(np.ceil(imshape[0] / Ts)), Ts * int(np.ceil(imshape[1] / Ts)))
    ly, lx = padded
    for lvl, (factor, ts) in enumerate(zip(bm.factors, bm.tile_sizes)):
        ly, lx = np.floor(ly / factor), np.floor(lx / factor)
        if ly / ts < 1 or lx / ts < 1:
            raise ValueError("Image of shape {} is incompatible with the given block matching tile sizes and "
                             "factors : at level {}, coarse image of shape {} cannot be divided into tiles of "
                             "size {}.".format(imshape, lvl, (ly, lx), ts))
    for lvl, ((h, w), ts) in enumerate(zip(pyramid_shapes(padded, bm.factors), bm.tile_sizes)):
        if h // ts < 1 or w // ts < 1:
            raise ValueError("Image of shape {} is incompatible with the block matching pyramid: level {} has "
                             "shape {} after the valid Gaussian filtering, smaller than one tile of size {}."
                             .format(imshape, lvl, (h, w), ts))
    valid = ["nearest", "bilinear", "bicubic"]
    assert bm.flow_upscale_mode in valid, \
        f"Unknown flow upscaling mode {bm.flow_upscale_mode}, should be one of {valid}."
