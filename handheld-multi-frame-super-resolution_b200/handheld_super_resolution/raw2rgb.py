"""Output side of process() on the DEVICE — mirrors handheld_super_resolution/raw2rgb.py of the reference
(get_color_matrix :118-136, apply_ccm :139-146, gamma_compression :143-146, devignette :203-210, postprocess :212-250)
plus the quantisation run_handheld.py applies before saving (:132-150).  The reference runs all of this in NumPy /
scikit-image on the host after a 576 MB float32 device-to-host copy; here the merged image stays on the GPU and only
the finished image — float32, or uint8 / uint16 on request — is copied back."""
import ctypes as C

import numpy as np
import torch

from . import _lib

RGB2XYZ = np.array([[0.4124564, 0.3575761, 0.1804375],
                    [0.2126729, 0.7151522, 0.0721750],
                    [0.0193339, 0.1191920, 0.9503041]])
OUT_KINDS = {None: (0, torch.float32), "float32": (0, torch.float32), "uint8": (1, torch.uint8), "uint16": (2, torch.uint16)}


def get_color_matrix(raw=None, xyz2cam=None):
    """rgb2cam (float32 3x3, rows normalised) from the camera's XYZ->cam matrix (raw2rgb.py:118-136).  `raw` is a rawpy
    object in the reference (only its rgb_xyz_matrix is read); pass xyz2cam directly when there is none."""
    if xyz2cam is None:
        if raw is None:
            raise ValueError("either a rawpy object or xyz2cam is needed for the colour matrix")
        xyz2cam = raw.rgb_xyz_matrix[:3]
    xyz2cam = np.asarray(xyz2cam)
    if np.linalg.norm(xyz2cam) == 0:
        print("Warning -- CCM not found or given. Use eye matrix instead.")
        rgb2cam = RGB2XYZ
    else:
        rgb2cam = xyz2cam @ RGB2XYZ
    return (rgb2cam / rgb2cam.sum(axis=-1, keepdims=True)).astype(np.float32)


def gaussian_taps(sigma, truncate=4.0):
    """(radius, float64 weights) of scipy.ndimage.gaussian_filter1d(sigma, truncate=4) — what skimage's unsharp_mask uses."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) ** 2) * x ** 2)
    return radius, np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)


def apply_ccm(image, ccm):
    """image [H,W,3] (CUDA) <- clip(ccm @ pixel, 0, 1), in place (raw2rgb.py:139-146 and the clip of :226)."""
    assert image.ndim == 3 and image.shape[-1] == 3
    m = np.ascontiguousarray(ccm, dtype=np.float32).reshape(9)
    _lib.call("hhsr_post_ccm_clip", _lib.ptr(image), image.shape[0] * image.shape[1], m.ctypes.data_as(C.POINTER(C.c_float)),
              _lib.stream())
    return image


def postprocess(raw, img=None, do_color_correction=True, do_tonemapping=True, do_gamma=True, sharpening_config=None,
                do_devignette=False, xyz2cam=None, output_dtype=None):
    """raw2rgb.postprocess (:212-250) of the merged linear image `img` [H,W,3] float32 — a CUDA tensor (or anything
    _lib.as_device accepts).  Returns a CUDA tensor [H,W,3]: float32 in [0,1] with NaN kept (what the reference returns),
    or, with output_dtype "uint8" / "uint16" (B200 addition), the quantised image run_handheld.py writes:
    rint(clip(nan_to_num(x), 0, 1) * 255 | 65535).  `img` is not modified unless colour correction is on.

    Not on the device: do_tonemapping (OpenCV's MergeMertens exposure fusion, raw2rgb.py:153-170) raises
    NotImplementedError; img=None (rawpy's own ISP, :217-219) needs rawpy."""
    if img is None:
        raise NotImplementedError("rawpy's whole-stack post-processing needs rawpy (raw2rgb.py:217-219)")
    if do_tonemapping:
        raise NotImplementedError("do_tonemapping uses OpenCV MergeMertens on the host (raw2rgb.py:153-170); not on the device path")
    if output_dtype not in OUT_KINDS:
        raise ValueError("output_dtype must be one of None, 'float32', 'uint8', 'uint16'")
    img = _lib.as_device(img)
    assert img.ndim == 3 and img.shape[-1] == 3
    H, W, _ = img.shape
    if do_color_correction:
        cam2rgb = np.linalg.inv(get_color_matrix(raw, xyz2cam))
        img = apply_ccm(img, cam2rgb)
    tmp, radius, taps, amount = None, 0, None, 0.0
    if sharpening_config is not None and sharpening_config.enabled:
        if "radius" in sharpening_config and "amount" in sharpening_config:
            sigma, amount = sharpening_config.radius, sharpening_config.amount
        else:
            import warnings
            warnings.warn("Sharpening config is missing radius or amount parameter, using default values.")
            sigma, amount = 3, 0.5
        radius, taps = gaussian_taps(sigma)
        tmp = torch.empty_like(img)
        _lib.call("hhsr_post_blur_cols", _lib.ptr(img), H, W, taps.ctypes.data_as(C.POINTER(C.c_double)), radius, _lib.ptr(tmp),
                  _lib.stream())
    kind, dtype = OUT_KINDS[output_dtype]
    out = torch.empty((H, W, 3), dtype=dtype, device=img.device)
    _lib.call("hhsr_post_finish", _lib.ptr(img), _lib.ptr(tmp), H, W,
              taps.ctypes.data_as(C.POINTER(C.c_double)) if taps is not None else None, radius, float(amount),
              int(bool(do_devignette)), float(np.float32(1.0 / 2.2)) if do_gamma else 0.0, kind, _lib.ptr(out), _lib.stream())
    return out
