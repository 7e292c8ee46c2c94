"""Coarse-to-fine alignment driver — mirrors handheld_super_resolution/alignment.py of the reference
(init_alignment :20-72, build_gaussian_pyramid :74-82, align :84-122, align_lvl :125-147, upscale_lvl :150-172).
No host synchronisation anywhere (the reference issues one cuda.synchronize() per level, alignment.py:105-107):
everything is enqueued on the current stream."""
import torch

from . import _lib
from .ICA import init_ica, align_lvl_ica
from .block_matching import align_lvl_block_matching_L2, align_lvl_block_matching_L1
from .utils_image import cuda_downsample

_MODES = {"nearest": 0, "bilinear": 1, "bicubic": 2}


def build_gaussian_pyramid(image, factors=[1, 2, 4, 4], kernel="gaussian"):
    """alignment.py:74-82 — levels returned coarse -> fine."""
    image = _lib.as_device(image)
    image = image.reshape(image.shape[-2], image.shape[-1])
    pyramid = [cuda_downsample(image, kernel, factors[0])]
    for factor in factors[1:]:
        pyramid.append(cuda_downsample(pyramid[-1], kernel, factor))
    return pyramid[::-1]


def init_alignment(ref_img, config):
    """Reference-side products (alignment.py:20-72): circular padding to a multiple of the tile size, pyramid,
    per-level gradients and tile Hessians.  Returns the reference's 6-tuple
    (pyramid, tiled_pyr, tiled_fft, gradx_pyr, grady_pyr, hessian_pyr), coarse -> fine; `tiled_pyr` holds the plain
    levels and `tiled_fft` holds [ny, nx] tile-grid shapes (opaque handles: only this module reads them)."""
    ref_img = _lib.as_device(ref_img)
    h, w = ref_img.shape
    ts0 = config.block_matching.tuning.tile_size
    tile_sizes = config.block_matching.tuning.tile_sizes
    hp = h + (ts0 - h % ts0) * (h % ts0 != 0)
    wp = w + (ts0 - w % ts0) * (w % ts0 != 0)
    if (hp, wp) != (h, w):
        padded = torch.empty((hp, wp), dtype=torch.float32, device=ref_img.device)
        _lib.call("hhsr_pad_circular", _lib.ptr(ref_img), h, w, _lib.ptr(padded), hp, wp, _lib.stream())
    else:
        padded = ref_img
    factors = config.block_matching.tuning.factors
    pyramid = build_gaussian_pyramid(padded, factors)
    tiled_fft, tiled_pyr, gradx_pyr, grady_pyr, hessian_pyr = [], [], [], [], []
    for i, lvl in enumerate(pyramid):
        ts = tile_sizes[len(factors) - i - 1]
        gradx, grady, hessian = init_ica(lvl, ts, config)
        if hessian.shape[0] < 1 or hessian.shape[1] < 1:
            raise ValueError("pyramid level %d of shape %s holds no tile of size %d (SURVEY Q12)"
                             % (len(factors) - i - 1, tuple(lvl.shape), ts))
        tiled_pyr.append(lvl)
        tiled_fft.append(tuple(hessian.shape[:2]))
        gradx_pyr.append(gradx), grady_pyr.append(grady), hessian_pyr.append(hessian)
    return pyramid, tiled_pyr, tiled_fft, gradx_pyr, grady_pyr, hessian_pyr


def align(ref_pyramid, tyled_pyr, ref_tiled_fft, ref_gradx, ref_grady, ref_hessian, img, config):
    """Flow of one moving grey image against the reference (alignment.py:84-122): [ny, nx, 2] float32 (dx, dy)."""
    factors = config.block_matching.tuning.factors
    moving_pyramid = build_gaussian_pyramid(img, factors)
    alignments = None
    for l, (ref_lvl, tyled_lvl, fft_lvl, gx, gy, hess, moving_lvl) in enumerate(zip(
            ref_pyramid, tyled_pyr, ref_tiled_fft, ref_gradx, ref_grady, ref_hessian, moving_pyramid)):
        list_id = len(ref_pyramid) - l - 1
        npatchs = tuple(hess.shape[:2])
        if alignments is None:
            alignments = torch.zeros((*npatchs, 2), dtype=torch.float32, device=ref_lvl.device)
        else:
            alignments = upscale_lvl(alignments, npatchs, list_id, config)
        align_lvl(ref_lvl, tyled_lvl, fft_lvl, gx, gy, hess, moving_lvl, alignments, l=list_id, config=config)
    return alignments


def align_lvl(ref_lvl, tyled_pyr_lvl, ref_fft_lvl, ref_gradx_lvl, ref_grady_lvl, ref_hessian_lvl,
              moving_lvl, alignments, l, config):
    """One pyramid level: block matching then ICA, both in place on `alignments` (alignment.py:125-147)."""
    metric = config.block_matching.tuning.metrics[l]
    if metric == "L2":
        align_lvl_block_matching_L2(tyled_pyr_lvl, ref_fft_lvl, moving_lvl, alignments, l, config)
    elif metric == "L1":
        align_lvl_block_matching_L1(ref_lvl, moving_lvl, alignments, l, config)
    else:
        raise ValueError("Unknown block matching metric {}".format(metric))
    align_lvl_ica(ref_lvl, ref_gradx_lvl, ref_grady_lvl, ref_hessian_lvl, moving_lvl, alignments, l, config)


def upscale_lvl(alignments, npatchs, l, config):
    """Re-tile and scale the flow for the next finer level (alignment.py:150-172)."""
    bm = config.block_matching.tuning
    new_ts, prev_ts, factor = bm.tile_sizes[l], bm.tile_sizes[l + 1], bm.factors[l + 1]
    repeat = factor // (new_ts // prev_ts)
    mode = bm.flow_upscale_mode
    if mode not in _MODES:
        raise ValueError("Unknown flow upscaling mode %s" % mode)
    ny_in, nx_in, _ = alignments.shape
    # the reference only pads when one side is too small; otherwise the upsampled grid keeps its own size
    ny_up, nx_up = ny_in * repeat, nx_in * repeat
    if ny_up < npatchs[0] or nx_up < npatchs[1]:
        ny_out, nx_out = npatchs
    else:
        ny_out, nx_out = ny_up, nx_up
    out = torch.empty((ny_out, nx_out, 2), dtype=torch.float32, device=alignments.device)
    _lib.call("hhsr_upscale_flow", _lib.ptr(alignments), ny_in, nx_in, _lib.ptr(out), ny_out, nx_out, int(repeat),
              float(factor), _MODES[mode], _lib.stream())
    return out
