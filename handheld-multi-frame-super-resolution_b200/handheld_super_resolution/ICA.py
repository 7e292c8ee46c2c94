"""Inverse-compositional Lucas-Kanade — mirrors handheld_super_resolution/ICA.py of the reference
(init_ica :15-34, align_lvl_ica :78-103)."""
import torch

from . import _lib


def init_ica(image, tile_size, config=None):
    """Gradients and per-tile Hessians of a reference pyramid level (ICA.py:15-34).
    Returns (gradx [h,w], grady [h,w], hessian [h//ts, w//ts, 2, 2])."""
    image = _lib.as_device(image)
    h, w = image.shape
    ny, nx = h // tile_size, w // tile_size
    gradx, grady = torch.empty_like(image), torch.empty_like(image)
    hessian = torch.empty((ny, nx, 2, 2), dtype=torch.float32, device=image.device)
    _lib.call("hhsr_grad_hessian", _lib.ptr(image), h, w, int(tile_size), _lib.ptr(gradx), _lib.ptr(grady),
              _lib.ptr(hessian), _lib.stream())
    return gradx, grady, hessian


def align_lvl_ica(ref_img, ref_gradx_lvl, ref_grady_lvl, ref_hessian_lvl, moving_lvl, alignment, l, config):
    """n_iter ICA iterations on every tile, updating `alignment` [ny,nx,2] IN PLACE (ICA.py:78-103)."""
    tile_size = config.block_matching.tuning.tile_sizes[l]
    if tile_size not in (8, 16, 32, 64):
        raise NotImplementedError("ICA kernel for tile size {} not implemented".format(tile_size))
    assert alignment.is_cuda and alignment.is_contiguous() and alignment.dtype == torch.float32
    ny, nx, _ = alignment.shape
    rh, rw = ref_img.shape
    mh, mw = moving_lvl.shape
    _lib.call("hhsr_ica", _lib.ptr(ref_img), _lib.ptr(ref_gradx_lvl), _lib.ptr(ref_grady_lvl), rh, rw,
              _lib.ptr(ref_hessian_lvl), _lib.ptr(moving_lvl), mh, mw, _lib.ptr(alignment), ny, nx, int(tile_size),
              int(config.ica.tuning.n_iter), _lib.stream())
