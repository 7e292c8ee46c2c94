"""Inverse-compositional Lucas-Kanade — mirrors handheld_super_resolution/ICA.py of the reference
(init_ica :15-34, align_lvl_ica :78-103)."""
import torch

from . import _lib


ON_THE_FLY_GRADIENTS = True     # False: always read the gradient planes (tests compare the two)


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
    # remember what these planes are: align_lvl_ica re-forms central differences inside its kernel instead of reading them
    gradx._hhsr_grad_of = grady._hhsr_grad_of = (image.data_ptr(), image._version, gradx.data_ptr(), grady.data_ptr())
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
    # gradients that init_ica derived from exactly this (unmodified) level need not be read back: the 32 x 32 kernel
    # re-forms them (identical values); anything else — user-supplied planes, other tile sizes, ragged levels — is read
    tag = (ref_img.data_ptr(), ref_img._version, ref_gradx_lvl.data_ptr(), ref_grady_lvl.data_ptr())
    if (ON_THE_FLY_GRADIENTS and tile_size == 32 and getattr(ref_gradx_lvl, "_hhsr_grad_of", None) == tag
            and getattr(ref_grady_lvl, "_hhsr_grad_of", None) == tag and ref_gradx_lvl._version == 0 and ref_grady_lvl._version == 0
            and ny * 32 == rh and nx * 32 == rw and rw % 4 == 0 and ref_img.data_ptr() % 16 == 0):
        ref_gradx_lvl = ref_grady_lvl = None
    _lib.call("hhsr_ica", _lib.ptr(ref_img), _lib.ptr(ref_gradx_lvl), _lib.ptr(ref_grady_lvl), rh, rw,
              _lib.ptr(ref_hessian_lvl), _lib.ptr(moving_lvl), mh, mw, _lib.ptr(alignment), ny, nx, int(tile_size),
              int(config.ica.tuning.n_iter), _lib.stream())
