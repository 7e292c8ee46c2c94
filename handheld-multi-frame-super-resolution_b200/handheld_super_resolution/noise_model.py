"""Noise curves sigma(b), d(b) for 1001 brightness levels, feeding the robustness stage.

Mirrors handheld_super_resolution/fast_monte_carlo.py:157-230 of the reference (run_fast_MC: Monte-Carlo on the
clipped ends of the brightness range, linear interpolation of sigma^2 and d^2 in between), with two changes: the
Monte-Carlo runs on the GPU (hhsr_noise_mc, one CTA per brightness level) instead of a multiprocessing pool, and it is
SEEDED (counter-based Philox) — the reference draws from the global numpy RNG in worker processes, so its `process()`
is not reproducible run to run.  regular_MC_numpy is the same estimator
in NumPy, kept for cross-checking the kernel in the tests; the product path never calls it."""
import numpy as np

N_PATCHES = int(1e5)
N_BRIGHTNESS_LEVELS = 1000
TOL = 3


def get_non_linearity_bound(alpha, beta, tol):
    """fast_monte_carlo.py:24-29."""
    tol_sq = tol * tol
    xmin = tol_sq / 2 * (alpha + np.sqrt(tol_sq * alpha * alpha + 4 * beta))
    xmax = (2 + tol_sq * alpha - np.sqrt((2 + tol_sq * alpha) ** 2 - 4 * (1 + tol_sq * beta))) / 2
    return xmin, xmax


def unitary_MC(alpha, beta, b, rng, n_patches=N_PATCHES):
    """fast_monte_carlo.py:31-66: mean |difference of 3x3 means| and mean 3x3 std of two noisy clipped patches."""
    std = np.sqrt(b * alpha + beta)
    p1 = np.clip(b + std * rng.standard_normal((n_patches, 9)), 0.0, 1.0)
    p2 = np.clip(b + std * rng.standard_normal((n_patches, 9)), 0.0, 1.0)
    std_mean = 0.5 * np.mean(np.std(p1, axis=1) + np.std(p2, axis=1))
    diff_mean = np.mean(np.abs(p1.mean(axis=1) - p2.mean(axis=1)))
    return diff_mean, std_mean


def regular_MC_numpy(b_array, alpha, beta, seed=0, n_patches=N_PATCHES):
    """fast_monte_carlo.py:68-101 in one NumPy process (test cross-check of the device kernel)."""
    rng = np.random.default_rng(seed)
    sigmas, diffs = np.empty_like(b_array), np.empty_like(b_array)
    for i, b in enumerate(b_array):
        diffs[i], sigmas[i] = unitary_MC(alpha, beta, b, rng, n_patches)
    return sigmas, diffs


def regular_MC(b_array, alpha, beta, seed=0, n_patches=N_PATCHES):
    """(sigmas, diffs) of the given brightness levels by Monte-Carlo on the GPU (fast_monte_carlo.py:68-101)."""
    import torch
    from . import _lib
    b = _lib.as_device(np.ascontiguousarray(b_array, dtype=np.float64), torch.float64)
    diffs, sigmas = torch.empty_like(b), torch.empty_like(b)
    _lib.call("hhsr_noise_mc", _lib.ptr(b), b.numel(), float(alpha), float(beta), int(n_patches), int(seed) & (2 ** 64 - 1),
              _lib.ptr(diffs), _lib.ptr(sigmas), _lib.stream())
    return sigmas.cpu().numpy(), diffs.cpu().numpy()


def run_fast_MC(alpha, beta, seed=0, n_patches=N_PATCHES, mc=None):
    """Returns (std_curve, diff_curve), 1001 float64 each (fast_monte_carlo.py:157-230).  `mc` replaces the Monte-Carlo
    estimator (signature of regular_MC; the tests pass regular_MC_numpy)."""
    mc = regular_MC if mc is None else mc
    n = N_BRIGHTNESS_LEVELS
    xmin, xmax = get_non_linearity_bound(alpha, beta, TOL)
    imin = int(np.ceil(xmin * n)) + 1
    imax = int(np.floor(xmax * n)) - 1
    brightness = np.arange(n + 1) / n
    if imin > n:
        return mc(brightness, alpha, beta, seed, n_patches)
    sigmas, diffs = np.empty(n + 1), np.empty(n + 1)
    nl = np.concatenate((brightness[:imin + 1], brightness[imax:]))
    s_nl, d_nl = mc(nl, alpha, beta, seed, n_patches)
    sigmas[:imin + 1], diffs[:imin + 1] = s_nl[:imin + 1], d_nl[:imin + 1]
    sigmas[imax:], diffs[imax:] = s_nl[imin + 1:], d_nl[imin + 1:]
    b_l = brightness[imin - 1:imax + 2]
    norm_b = (b_l - b_l[0]) / (b_l[-1] - b_l[0])
    sigmas[imin:imax + 1] = np.sqrt(norm_b * (sigmas[imax] ** 2 - sigmas[imin] ** 2) + sigmas[imin] ** 2)[1:-1]
    diffs[imin:imax + 1] = np.sqrt(norm_b * (diffs[imax] ** 2 - diffs[imin] ** 2) + diffs[imin] ** 2)[1:-1]
    return sigmas, diffs
