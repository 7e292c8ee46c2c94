"""Image helpers on the hot path — mirrors handheld_super_resolution/utils_image.py of the reference:
compute_grey_images (:58-115), GAT (:117-170), cuda_downsample (:360-391)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


_GREY_PLANS = {}      # (device, h, w) -> (plan tables on the device, bytes of the per-call work buffer) or None
GREY_FFT_NATIVE = True   # False: always take the cuFFT route (tests compare the two)


def _grey_plan(h, w, device):
    """Twiddle / digit-reversal tables of the library's own grey-image FFT for an h x w frame, built once per shape and
    device; None when a size has a prime factor above 19 (those frames take the cuFFT route)."""
    key = (device.index, h, w)
    if key not in _GREY_PLANS:
        L = _lib.lib()
        pb, wb = C.c_size_t(), C.c_size_t()
        rc = L.hhsr_grey_fft_sizes(h, w, C.byref(pb), C.byref(wb))
        if rc == -2:            # HHSR_E_UNSUPPORTED
            _GREY_PLANS[key] = None
        elif rc != 0:
            raise RuntimeError("hhsr_grey_fft_sizes failed (%d): %s" % (rc, L.hhsr_last_error_string().decode()))
        else:
            plan = torch.empty(pb.value, dtype=torch.uint8, device=device)
            _lib.call("hhsr_grey_fft_plan", _lib.ptr(plan), h, w, _lib.stream())
            torch.cuda.current_stream(device).synchronize()     # once per shape: other streams read the tables later
            _GREY_PLANS[key] = (plan, wb.value)
    return _GREY_PLANS[key]


def compute_grey_images(img, method):
    """Raw -> grey (utils_image.py:58-115).

    "FFT": the ideal half-band low-pass of Alg. 3.  Frames whose sizes factor into primes up to 19 (4000 x 3000, 4032 x 3024, 5472 x 3648, 8192 x 6144, ...) go
    through the library's own three shared-memory FFT passes (hhsr_grey_fft: rows forward, columns forward + band mask +
    inverse, rows inverse; only the quarter of the spectrum the mask keeps is ever stored).  Other sizes: cuFFT
    (torch.fft.rfft2 / irfft2) around the in-place band-mask kernel hhsr_grey_band_mask — the reference's fftshift +
    four masked fills + ifftshift + .real on the half spectrum, mathematically identical for a real input.
    "decimating": 2x2 mean."""
    img = _lib.as_device(img)
    h, w = img.shape
    if method == "FFT":
        plan = _grey_plan(h, w, img.device) if GREY_FFT_NATIVE else None
        if plan is not None:
            work = torch.empty(plan[1], dtype=torch.uint8, device=img.device)
            out = torch.empty_like(img)
            _lib.call("hhsr_grey_fft", _lib.ptr(img), h, w, _lib.ptr(plan[0]), _lib.ptr(work), _lib.ptr(out), _lib.stream())
            return out
        spec = torch.fft.rfft2(img)
        # the 1/(h*w) of the inverse transform is folded into the mask (kept entries are scaled, the rest zeroed) and
        # the inverse runs unnormalised: one full-image multiply less per frame
        _lib.call("hhsr_grey_band_mask", _lib.ptr(spec), h, w, spec.stride(0), spec.stride(1), 1.0 / (h * w), _lib.stream())
        return torch.fft.irfft2(spec, s=(h, w), norm="forward")
    elif method == "decimating":
        out = torch.empty((h // 2, w // 2), dtype=torch.float32, device=img.device)
        _lib.call("hhsr_decimate_to_grey", _lib.ptr(img), h, w, _lib.ptr(out), _lib.stream())
        return out
    raise NotImplementedError("Computation of gray level on GPU is only supported for FFT")


def GAT(image, alpha, beta):
    """Generalised Anscombe transform (utils_image.py:117-170)."""
    image = _lib.as_device(image)
    assert len(image.shape) == 2
    assert alpha > 0, f"alpha should be positive, got {alpha} (VST is ill defined and kernels would be wrong)"
    out = torch.empty_like(image)
    _lib.call("hhsr_gat", _lib.ptr(image), image.numel(), float(alpha), float(beta), _lib.ptr(out), _lib.stream())
    return out


def gaussian_kernel1d(sigma, radius):
    """The 1-D kernel the reference takes from scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius)
    (utils_image.py:380), restated so that no private scipy API is needed."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


_TAPS = {}


def _downsample_taps(factor):
    """(radius, float32 taps) of one pyramid level, cached per factor (built per frame and level otherwise)."""
    if factor not in _TAPS:
        radius = int(4 * factor * 0.5 + 0.5)
        _TAPS[factor] = (radius, np.ascontiguousarray(gaussian_kernel1d(factor * 0.5, radius)[::-1].astype(np.float32)))
    return _TAPS[factor]


def cuda_downsample(th_img, kernel="gaussian", factor=2):
    """One pyramid level (utils_image.py:360-391).  th_img: [h,w] (or [1,1,h,w]) CUDA tensor; returns [h2,w2]."""
    if factor == 1:
        return th_img
    if kernel != "gaussian":
        raise ValueError("please use gaussian kernel")
    img = _lib.as_device(th_img)
    img = img.reshape(img.shape[-2], img.shape[-1])
    radius, taps = _downsample_taps(int(factor))
    h, w = img.shape
    h2, w2 = (h - 2 * radius) // factor, (w - 2 * radius) // factor
    if h2 < 1 or w2 < 1:
        raise ValueError("image of shape %s too small for a pyramid level of factor %d" % ((h, w), factor))
    out = torch.empty((h2, w2), dtype=torch.float32, device=img.device)
    _lib.call("hhsr_gauss_downsample", _lib.ptr(img), h, w, int(factor),
              taps.ctypes.data_as(C.POINTER(C.c_float)), radius, _lib.ptr(out), h2, w2, _lib.stream())
    return out


def apply_orientation(img, ori):
    """EXIF orientation of the output image (utils_image.py:12-56), numpy views on the host."""
    if ori == 2:
        img = np.flip(img, axis=1)
    elif ori == 3:
        img = np.rot90(img, k=2, axes=(0, 1))
    elif ori == 4:
        img = np.flip(img, axis=0)
    elif ori == 5:
        img = np.rot90(np.flip(img, axis=1), k=-3, axes=(0, 1))
    elif ori == 6:
        img = np.rot90(img, k=-1, axes=(0, 1))
    elif ori == 7:
        img = np.rot90(np.flip(img, axis=1), k=-1, axes=(0, 1))
    elif ori == 8:
        img = np.rot90(img, k=-3, axes=(0, 1))
    return img


def _frame_count_denoise(entry, image, r_acc, strength, max_frame_count, scale):
    image = _lib.as_device(image)
    r_acc = _lib.as_device(r_acc, torch.float64)
    assert image.ndim == 3 and image.shape[-1] == 3 and r_acc.ndim == 2
    denoised = torch.empty_like(image)
    _lib.call(entry, _lib.ptr(image), image.shape[0], image.shape[1], _lib.ptr(r_acc), r_acc.shape[0], r_acc.shape[1],
              float(scale), float(strength), float(max_frame_count), _lib.ptr(denoised), _lib.stream())
    return denoised


def frame_count_denoising_gauss(image, r_acc, config, scale=None, mode="bayer"):
    """Gaussian blur of the merged image whose sigma grows where few frames were accumulated (utils_image.py:174-231).
    `config` is the `accumulated_robustness_denoiser.gauss` node (sigma_max, max_frame_count).  Upstream this stage
    cannot run: it reads `config.mode` / `config.scale` from that node (:177-178) and iterates range() over a float
    (:210-215); here mode / scale are arguments (process() passes the main configuration's) and the window half-width is
    ceil(3 sigma).  Returns a new CUDA tensor."""
    if config.get("mode", mode) != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    scale = config.get("scale", scale)
    if scale is None:
        raise ValueError("frame_count_denoising_gauss needs the scale of the merge (config.scale)")
    return _frame_count_denoise("hhsr_frame_count_denoise_gauss", image, r_acc, config.sigma_max, config.max_frame_count, scale)


def frame_count_denoising_median(image, r_acc, config, scale=None, mode="bayer"):
    """Median filter of the merged image whose radius grows where few frames were accumulated (utils_image.py:233-315):
    literal bubble sort and upper median like the reference kernel; radius_max <= 7 (the reference's 256-entry window
    buffer overflows above).  Same remarks on `config` / scale / mode as frame_count_denoising_gauss."""
    if config.get("mode", mode) != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    scale = config.get("scale", scale)
    if scale is None:
        raise ValueError("frame_count_denoising_median needs the scale of the merge (config.scale)")
    return _frame_count_denoise("hhsr_frame_count_denoise_median", image, r_acc, config.radius_max, config.max_frame_count, scale)


def computeRMSE(image1, image2):
    """utils_image.py:408-414."""
    assert np.array_equal(image1.shape, image2.shape), "images have different sizes"
    d = image1.astype(np.float64) - image2.astype(np.float64)
    return np.sqrt(np.mean(d * d))


def computePSNR(image, noisyImage):
    """utils_image.py:417-437 (images in [0,1])."""
    rmse = computeRMSE(image, noisyImage)
    return float("inf") if rmse == 0 else 20 * np.log10(1.0 / rmse)
