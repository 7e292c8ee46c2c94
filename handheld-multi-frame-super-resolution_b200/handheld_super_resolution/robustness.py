"""Robustness estimation (Alg. 6-9) — mirrors handheld_super_resolution/robustness.py of the reference
(init_robustness :23-76, compute_robustness :79-170 and the sub-stages it exposes)."""
import torch

from . import _lib


def compute_guide_stats(raw_img, cfa_pattern, white_balance, need_vars=True):
    """Guide image (Alg. 7, robustness.py:173-226) fused with its 3x3 local statistics (Alg. 8, :228-294).
    Returns (means, vars) [3, H//2, W//2]; vars is None when not requested."""
    raw_img = _lib.as_device(raw_img)
    H, W = raw_img.shape
    means = torch.empty((3, H // 2, W // 2), dtype=torch.float32, device=raw_img.device)
    vars_ = torch.empty_like(means) if need_vars else None
    _lib.call("hhsr_guide_stats", _lib.ptr(raw_img), H, W, _lib.cfa_array(cfa_pattern), _lib.wb_array(white_balance),
              _lib.ptr(means), _lib.ptr(vars_), _lib.stream())
    return means, vars_


def compute_guide_image(raw_img, cfa_pattern, white_balance):
    """Guide image G [3, H//2, W//2] of a raw frame (Alg. 7, robustness.py:173-226) — the stand-alone stage; the
    pipeline uses compute_guide_stats, which fuses it with compute_local_stats (bit-equal)."""
    raw_img = _lib.as_device(raw_img)
    H, W = raw_img.shape
    guide = torch.empty((3, H // 2, W // 2), dtype=torch.float32, device=raw_img.device)
    _lib.call("hhsr_guide_image", _lib.ptr(raw_img), H, W, _lib.cfa_array(cfa_pattern), _lib.wb_array(white_balance),
              _lib.ptr(guide), _lib.stream())
    return guide


def compute_local_stats(guide_img):
    """3x3 local mean and variance of the guide image (Alg. 8, robustness.py:228-294): (means, vars) [C, h, w]."""
    guide_img = _lib.as_device(guide_img)
    n_channels, h, w = guide_img.shape
    if n_channels not in (1, 3):
        raise ValueError("Incoherent number of channel : {}".format(n_channels))
    means, vars_ = torch.empty_like(guide_img), torch.empty_like(guide_img)
    _lib.call("hhsr_local_stats", _lib.ptr(guide_img), n_channels, h, w, _lib.ptr(means), _lib.ptr(vars_), _lib.stream())
    return means, vars_


def upscale_warp_stats(local_stats, tile_size=None, flow=None):
    """x2 Dodgson upsampling (+ warp by the tile flow) of a [3,h,w] statistic to [3,2h,2w] (robustness.py:296-418)."""
    local_stats = _lib.as_device(local_stats)
    c, h, w = local_stats.shape
    if c != 3:
        raise ValueError("Incoherent number of channel : {}".format(c))
    out = torch.empty((3, 2 * h, 2 * w), dtype=torch.float32, device=local_stats.device)
    if flow is None:
        _lib.call("hhsr_upscale_warp_stats", _lib.ptr(local_stats), h, w, None, 0, 0, 0, _lib.ptr(out), _lib.stream())
    else:
        flow = _lib.as_device(flow)
        _lib.call("hhsr_upscale_warp_stats", _lib.ptr(local_stats), h, w, _lib.ptr(flow), flow.shape[0], flow.shape[1],
                  int(tile_size), _lib.ptr(out), _lib.stream())
    return out


def init_robustness(ref_img, cfa_pattern, white_balance, config, noise_model=None, need_stds=True):
    """Local statistics of the reference frame, upsampled to raw resolution (robustness.py:23-76).
    Returns (local_means, local_stds) [3, H, W] — `local_stds` holds variances, like the reference.

    B200 addition: when `noise_model` (the curves or a table from noise_table()) is given, the upsampling of both
    statistics and the reference-side noise terms (see ref_noise_terms) come from ONE launch and the terms are cached
    on `local_means` for compute_robustness; with need_stds=False the upsampled variances are not even written
    (local_stds is None; main() uses this)."""
    if not config.robustness.enabled:
        return None, None
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    means, vars_ = compute_guide_stats(ref_img, cfa_pattern, white_balance)
    if noise_model is None:
        return upscale_warp_stats(means), upscale_warp_stats(vars_)
    table = noise_table(noise_model)
    _, h, w = means.shape
    up_means = torch.empty((3, 2 * h, 2 * w), dtype=torch.float32, device=means.device)
    up_vars = torch.empty_like(up_means) if need_stds else None
    terms = torch.empty((4, 2 * h, 2 * w), dtype=torch.float32, device=means.device)
    _lib.call("hhsr_ref_stats_terms", _lib.ptr(means), _lib.ptr(vars_), h, w, _lib.ptr(table), table.shape[0],
              _lib.ptr(up_means), _lib.ptr(up_vars), _lib.ptr(terms), _lib.stream())
    up_means._hhsr_noise_terms = (("fused", table.data_ptr()), terms, table)
    return up_means, up_vars


def noise_table(noise_model):
    """(std_curve, diff_curve) -> float32 device table [n, 2] of (sigma_t^2, d_t^2).  Accepts an already built table
    (main() builds it once per burst)."""
    if isinstance(noise_model, torch.Tensor):
        return noise_model
    std_curve, diff_curve = noise_model
    std_curve = _lib.as_device(std_curve, torch.float64)
    diff_curve = _lib.as_device(diff_curve, torch.float64)
    assert std_curve.numel() == diff_curve.numel()
    table = torch.empty((std_curve.numel(), 2), dtype=torch.float32, device=std_curve.device)
    _lib.call("hhsr_noise_table", _lib.ptr(std_curve), _lib.ptr(diff_curve), std_curve.numel(), _lib.ptr(table), _lib.stream())
    return table


def ref_noise_terms(ref_local_means, ref_local_stds, table):
    """Reference-side part of the noise model (robustness.py:504-533): [4, H, W] = (d_t^2 per channel, sum of
    max(sigma_p^2, sigma_t^2)).  Depends on the reference frame only, so it is built once per burst and cached on the
    `ref_local_means` tensor (the reference recomputes it for every comp frame)."""
    cached = getattr(ref_local_means, "_hhsr_noise_terms", None)
    if cached is not None and cached[0] == ("fused", table.data_ptr()):      # built by init_robustness(noise_model=...)
        return cached[1]
    if ref_local_stds is None:
        raise ValueError("ref_local_stds is required unless init_robustness() was given the noise model")
    ref_local_stds = _lib.as_device(ref_local_stds)
    key = (ref_local_stds.data_ptr(), table.data_ptr(), ref_local_means._version, ref_local_stds._version)
    if cached is not None and cached[0] == key:
        return cached[1]
    _, H, W = ref_local_means.shape
    terms = torch.empty((4, H, W), dtype=torch.float32, device=ref_local_means.device)
    _lib.call("hhsr_robustness_ref_terms", _lib.ptr(ref_local_means), _lib.ptr(ref_local_stds), H, W, _lib.ptr(table),
              table.shape[0], _lib.ptr(terms), _lib.stream())
    ref_local_means._hhsr_noise_terms = (key, terms, ref_local_stds, table)   # keep the keyed tensors alive
    return terms


def local_min(R, acc_rob=None, out=None):
    """5x5 local minimum (Alg. 9, robustness.py:641-687); optionally fused with `acc_rob += r` (utils.add).
    `out`: caller-owned result buffer (e.g. a symmetric-memory slot of the row-sharded multi-GPU merge)."""
    R = _lib.as_device(R)
    r = torch.empty_like(R) if out is None else out
    assert r.shape == R.shape and r.dtype == torch.float32 and r.is_contiguous() and r.data_ptr() != R.data_ptr()
    _lib.call("hhsr_local_min5", _lib.ptr(R), R.shape[0], R.shape[1], _lib.ptr(r), _lib.ptr(acc_rob), _lib.stream())
    return r


def compute_robustness(comp_img, ref_local_means, ref_local_stds, flows, cfa_pattern, white_balance, noise_model,
                       config, acc_rob=None, return_R=False, generic=False, out=None):
    """Robustness map r [H, W] of comp frame J_n (Alg. 6, robustness.py:79-170).  Three launches: guide statistics
    at half resolution, the fused per-pixel kernel (warp, distance, noise model, S, threshold), 5x5 minimum (plus,
    for the first comp frame of a burst, the reference-side noise terms).
    `acc_rob` (float64 [H,W]), when given, is incremented by r in the last launch (super_resolution.py:159);
    generic=True forces the per-pixel path of the fused kernel (A/B parity tests)."""
    comp_img = _lib.as_device(comp_img)
    H, W = comp_img.shape
    if not config.robustness.enabled:
        if out is not None:
            return out.fill_(1.0)
        return torch.ones((H, W), dtype=torch.float32, device=comp_img.device)
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    tun = config.robustness.tuning
    ts = config.block_matching.tuning.tile_size
    table = noise_table(noise_model)
    flows = _lib.as_device(flows)
    comp_means, _ = compute_guide_stats(comp_img, cfa_pattern, white_balance, need_vars=False)
    R = torch.empty((H, W), dtype=torch.float32, device=comp_img.device)
    ref_local_means = _lib.as_device(ref_local_means)
    terms = ref_noise_terms(ref_local_means, ref_local_stds, table)
    _lib.call("hhsr_robustness", _lib.ptr(comp_means), _lib.ptr(ref_local_means), _lib.ptr(terms), H, W,
              _lib.ptr(flows), flows.shape[0], flows.shape[1], int(ts),
              float(tun.t), float(tun.s1), float(tun.s2), float(tun.Mt), _lib.ptr(R), int(bool(generic)), _lib.stream())
    r = local_min(R, acc_rob, out=out)
    return (r, R) if return_R else r
