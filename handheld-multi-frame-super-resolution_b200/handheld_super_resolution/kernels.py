"""Steering-kernel estimation (Alg. 5) — mirrors handheld_super_resolution/kernels.py of the reference
(estimate_kernels :29-136)."""
import torch

from . import _lib

SEL_HARD_THRESHOLD = 0
SEL_LINEAR = 1


def estimate_kernels(img, config, out=None):
    """Covariance matrices Omega [H//2, W//2, 2, 2] of the merge kernels of raw frame `img` (kernels.py:29-136).
    One fused launch (GAT, decimation, gradients, structure tensor, eigen-decomposition, k1/k2, covariance)."""
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    mt = config.merging.tuning
    if config.merging.selection_law == "hard_threshold":
        law = SEL_HARD_THRESHOLD
    elif config.merging.selection_law == "linear":
        law = SEL_LINEAR
    else:
        raise ValueError(f"Unknown selection law: {config.merging.selection_law}")
    alpha, beta = config.noise_model.alpha, config.noise_model.beta
    assert alpha > 0, f"alpha should be positive, got {alpha} (VST is ill defined and kernels would be wrong)"
    img = _lib.as_device(img)
    H, W = img.shape
    covs = torch.empty((H // 2, W // 2, 2, 2), dtype=torch.float32, device=img.device) if out is None else out
    assert tuple(covs.shape) == (H // 2, W // 2, 2, 2) and covs.dtype == torch.float32 and covs.is_contiguous()
    _lib.call("hhsr_estimate_kernels", _lib.ptr(img), H, W, float(alpha), float(beta), float(mt.k_detail),
              float(mt.k_denoise), float(mt.D_th), float(mt.D_tr), float(mt.k_stretch), float(mt.k_shrink), law,
              _lib.ptr(covs), _lib.stream())
    return covs
