"""Shared test helpers: golden fixtures, config construction, NaN/inf-aware comparisons."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFA = [[0, 1], [1, 2]]
WB = [2.0, 1.0, 1.5, 0.0]


def load(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def curves():
    z = load("noise_curves_iso100.npz")
    return z["std_curve"], z["diff_curve"]


def plain_cfg(summary=None, **over):
    """Nested dict config (what the oracle takes) from a golden cfg summary (see make_golden_gpu.cfg_summary)."""
    c = dict(scale=2, tile_size=32, tile_sizes=[32, 32, 16], factors=[1, 2, 2], search_radii=[1, 4, 4],
             metrics=["L1", "L2", "L2"], n_iter=3, k_detail=0.25, k_denoise=3.0, D_th=0.71, D_tr=1.0, k_stretch=4,
             k_shrink=2, kernel="steerable", selection_law="linear", t=0.12, s1=2, s2=12, Mt=0.8,
             alpha=1.80710882e-4, beta=3.1937599182128e-6)
    if summary is not None:
        c.update(json.loads(str(summary)) if not isinstance(summary, dict) else summary)
    c.update(over)
    std, diff = curves()
    return {
        "scale": c["scale"], "mode": "bayer", "debug": False, "verbose": 0, "grey_method": "FFT",
        "block_matching": {"tuning": {"tile_size": c["tile_size"], "tile_sizes": list(c["tile_sizes"]),
                                      "factors": list(c["factors"]), "search_radii": list(c["search_radii"]),
                                      "metrics": list(c["metrics"]), "flow_upscale_mode": c.get("flow_upscale_mode", "nearest"),
                                      "tile_size_factors": [1] * (len(c["factors"]) - 1) + [0.5]}},
        "ica": {"tuning": {"n_iter": c["n_iter"], "sigma_blur": 0}},
        "merging": {"kernel": c["kernel"], "selection_law": c["selection_law"],
                    "tuning": {k: c[k] for k in ["k_detail", "k_denoise", "D_th", "D_tr", "k_stretch", "k_shrink"]}},
        "robustness": {"enabled": c.get("robustness_enabled", True), "save_mask": c.get("robustness_enabled", True),
                       "tuning": {k: c[k] for k in ["t", "s1", "s2", "Mt"]}},
        "noise_model": {"alpha": c["alpha"], "beta": c["beta"], "std_curve": std, "diff_curve": diff},
        "exif": {"cfa_pattern": CFA, "white_balance": WB, "iso": 100},
        "accumulated_robustness_denoiser": {"enabled": False,
                                            "median": {"enabled": False, "radius_max": 3, "max_frame_count": 8},
                                            "gauss": {"enabled": False, "sigma_max": 1.5, "max_frame_count": 8},
                                            "merge": {"enabled": False, "rad_max": 2, "max_multiplier": 8,
                                                      "max_frame_count": 2}},
    }


def attr_cfg(summary=None, **over):
    """Same config as an attribute-style Config (what the product takes)."""
    from handheld_super_resolution.config import Config
    return Config.wrap(plain_cfg(summary, **over))


def maxdiff(a, b):
    """max |a-b| over entries finite in both; asserts identical NaN and inf patterns."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    assert np.array_equal(np.isinf(a), np.isinf(b)), "inf pattern differs"
    m = np.isfinite(a) & np.isfinite(b)
    return float(np.abs(a[m] - b[m]).max()) if m.any() else 0.0


def reldiff(a, b, floor=1e-3):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    return float((np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), floor)).max()) if m.any() else 0.0
