"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's device pipeline (Jamy-L/Handheld-Multi-Frame-Super-Resolution,
`handheld_super_resolution/super_resolution.py:41-200` and the stage files it calls).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module; the
product (`handheld-multi-frame-super-resolution_b200/`) never does and fails loudly when its CUDA library is
missing.

Every function cites the reference file:line it follows (paths relative to the reference's
`handheld_super_resolution/`).  Arithmetic types follow what *compiled* Numba infers on the GPU (int64 op
float32 -> float64, Python-float kernel arguments are float64, `x[i] += y` rounds to the array dtype at every
step), not what the Python simulator does — see SURVEY.md Q8/Q9.

Parity status: PINNED.  The oracle is checked (tests/test_oracle_golden.py) against
  * tests/golden/{tiny_pipeline,medium_pipeline,stage_cases,alignment_cases}.npz — outputs of the unmodified
    reference run through real Numba-CUDA on a B200 (generator: tests/golden/make_golden_gpu.py);
  * tests/golden/cudasim_cases.npz — outputs of the unmodified reference kernels under NUMBA_ENABLE_CUDASIM in
    the build container (generator: tests/golden/make_golden_cudasim.py).
The reference itself ships no tests or golden vectors.
"""
import math

import numpy as np

F32 = np.float32
F64 = np.float64
EPSILON_DIV = 1e-10  # utils.py:21


def fma32(a, b, c):
    """float32 fused multiply-add (one rounding).  NVVM contracts `x*y + z` in the reference's compiled kernels;
    where that changes results visibly (structure tensor / eigen solve, kernels.py:138-191) the oracle restates
    the contraction the B200 goldens exhibit (found by matching tests/golden/tiny_pipeline.npz to 1 ulp)."""
    return (np.asarray(a, F64) * np.asarray(b, F64) + np.asarray(c, F64)).astype(F32)


def normalize_raw(raw, cfa, black_levels, white_level, white_balance):
    """utils_dng.py:146-160, literally: integer sensor counts [..., H, W] -> float32, per CFA position
    (x - black[c]) / (white - black[c]) then *= wb[c] / wb[1], with the reference's Python-scalar operands (ints for
    the levels, a float for the gain) so NumPy rounds them exactly as it does there."""
    x = np.asarray(raw).astype(F32)
    black = [int(b) for b in black_levels]
    white = int(white_level)
    wb = [float(w) for w in white_balance]
    for i in range(2):
        for j in range(2):
            channel = int(cfa[i][j])
            k = wb[channel] / wb[1]
            x[..., i::2, j::2] = (x[..., i::2, j::2] - black[channel]) / (white - black[channel])
            x[..., i::2, j::2] *= k
    return x


# --------------------------------------------------------------------------------------------------------------
# Grey image (Alg. 3) — utils_image.py:82-100
# --------------------------------------------------------------------------------------------------------------
def grey_fft(img):
    """Ideal low-pass keeping the centre half band of the (shifted) spectrum; utils_image.py:82-100."""
    h, w = img.shape
    f = np.fft.fftshift(np.fft.fft2(img.astype(F64)))
    f[:h // 4, :] = 0
    f[:, :w // 4] = 0
    f[-h // 4:, :] = 0          # -h//4 == -ceil(h/4)
    f[:, -w // 4:] = 0
    return np.fft.ifft2(np.fft.ifftshift(f)).real.astype(F32)


def grey_band_weights(n):
    """Per-axis keep flags of grey_fft on UNSHIFTED frequency indices: (M(k), M(-k)) for k = 0..n-1."""
    k = np.arange(n)
    s = (k + n // 2) % n
    keep = (s >= n // 4) & (s < n - (-(-n // 4)))
    return keep, keep[(-k) % n]


def decimate_to_grey(img):
    """utils_image.py:346-357 — `c = 0; c += img[...]` is int64+float32 -> float64; result stored f32."""
    h, w = img.shape[0] // 2, img.shape[1] // 2
    q = img[:2 * h, :2 * w].astype(F64)
    c = ((q[0::2, 0::2] + q[0::2, 1::2]) + q[1::2, 0::2]) + q[1::2, 1::2]
    return (c / 4).astype(F32)


def gat(img, alpha, beta):
    """Generalised Anscombe transform, utils_image.py:156-170 (alpha, beta are float64 kernel arguments)."""
    v = alpha * img.astype(F64) + 3 / 8 * alpha * alpha + beta
    v = np.maximum(0, v)
    return (2 / alpha * np.sqrt(v)).astype(F32)


# --------------------------------------------------------------------------------------------------------------
# Gaussian pyramid — alignment.py:74-82, utils_image.py:360-391
# --------------------------------------------------------------------------------------------------------------
def gaussian_kernel1d(sigma, radius):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) (scipy 1.18.1), used at utils_image.py:380."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def downsample(img, factor):
    """cuda_downsample, utils_image.py:360-391: valid separable correlation (y then x), stride-`factor` subsample."""
    if factor == 1:
        return img
    radius = int(4 * factor * 0.5 + 0.5)
    g = gaussian_kernel1d(factor * 0.5, radius)[::-1].astype(F32)
    K = 2 * radius + 1
    h, w = img.shape
    hv, wv = h - K + 1, w - K + 1
    h2, w2 = hv // factor, wv // factor
    rows = np.arange(h2) * factor
    tmp = np.zeros((h2, w), F32)
    for a in range(K):
        tmp += g[a] * img[rows + a, :]
    cols = np.arange(w2) * factor
    out = np.zeros((h2, w2), F32)
    for b in range(K):
        out += g[b] * tmp[:, cols + b]
    return out


def build_gaussian_pyramid(img, factors):
    """alignment.py:74-82 — returned coarse -> fine."""
    pyr = [downsample(img, factors[0])]
    for f in factors[1:]:
        pyr.append(downsample(pyr[-1], f))
    return pyr[::-1]


# --------------------------------------------------------------------------------------------------------------
# ICA init — ICA.py:15-76
# --------------------------------------------------------------------------------------------------------------
def gradients(img):
    """ICA.py:20-21: cross-correlation with [-1,0,1], zero 'same' padding."""
    gx = np.zeros_like(img)
    gy = np.zeros_like(img)
    gx[:, 1:-1] = img[:, 2:] - img[:, :-2]
    gx[:, 0] = img[:, 1]
    gx[:, -1] = -img[:, -2]
    gy[1:-1, :] = img[2:, :] - img[:-2, :]
    gy[0, :] = img[1, :]
    gy[-1, :] = -img[-2, :]
    return gx, gy


def tile_view(a, ts, ny, nx):
    return a[:ny * ts, :nx * ts].reshape(ny, ts, nx, ts).transpose(0, 2, 1, 3)


def compute_hessian(gx, gy, ts):
    """ICA.py:36-76: per-tile f32 sums accumulated sequentially in raster order."""
    ny, nx = gx.shape[0] // ts, gx.shape[1] // ts
    tx = tile_view(gx, ts, ny, nx).reshape(ny, nx, -1)
    ty = tile_view(gy, ts, ny, nx).reshape(ny, nx, -1)
    H = np.empty((ny, nx, 2, 2), F32)
    H[..., 0, 0] = np.cumsum(tx * tx, axis=-1, dtype=F32)[..., -1]
    H[..., 0, 1] = H[..., 1, 0] = np.cumsum(tx * ty, axis=-1, dtype=F32)[..., -1]
    H[..., 1, 1] = np.cumsum(ty * ty, axis=-1, dtype=F32)[..., -1]
    return H


def init_ica(img, ts):
    gx, gy = gradients(img)
    return gx, gy, compute_hessian(gx, gy, ts)


# --------------------------------------------------------------------------------------------------------------
# Block matching — block_matching.py
# --------------------------------------------------------------------------------------------------------------
def bm_l2_errors(ref, mov, flow, ts, r):
    """E[ty,tx,v,u] = sum m^2 - 2 sum ref*m over the tile, m read at clamped coordinates displaced by
    rint(flow)+(u,v); what block_matching.py:20-76 + extract_flow_patches :348-378 compute through FFTs."""
    ny, nx = flow.shape[:2]
    hm, wm = mov.shape
    f = np.rint(flow).astype(np.int64)
    reft = tile_view(ref.astype(F64), ts, ny, nx)
    E = np.empty((ny, nx, 2 * r + 1, 2 * r + 1), F64)
    ys = (np.arange(ny) * ts)[:, None, None, None] + np.arange(ts)[None, None, :, None]
    xs = (np.arange(nx) * ts)[None, :, None, None] + np.arange(ts)[None, None, None, :]
    movd = mov.astype(F64)
    for v in range(-r, r + 1):
        yy = np.clip(ys + f[..., 1][:, :, None, None] + v, 0, hm - 1)
        for u in range(-r, r + 1):
            xx = np.clip(xs + f[..., 0][:, :, None, None] + u, 0, wm - 1)
            m = movd[yy, xx]
            E[:, :, v + r, u + r] = (m * m).sum((-2, -1)) - 2 * (reft * m).sum((-2, -1))
    return E


def bm_l2(ref, mov, flow, ts, r, return_margin=False):
    """align_lvl_block_matching_L2: flow += first argmin (v-major).  Returns new flow (fraction kept)."""
    E = bm_l2_errors(ref, mov, flow, ts, r)
    n = 2 * r + 1
    Ef = E.reshape(*E.shape[:2], n * n)
    idx = np.argmin(Ef, axis=-1)
    out = flow.copy()
    out[..., 0] += (idx % n - r).astype(F32)
    out[..., 1] += (idx // n - r).astype(F32)
    if return_margin:
        srt = np.sort(Ef, axis=-1)
        scale = np.maximum(np.abs(srt[..., 0]), 1e-12)
        return out, (srt[..., 1] - srt[..., 0]) / scale
    return out


def bm_l1_compiled(flow):
    """What the compiled reference does at an L1 level with tile size 32/64 (confirmed on B200 by
    baseline/probe_reference.py): block_matching.py:241-252 never updates the shift, so flow <- rint(flow)."""
    return np.rint(flow).astype(F32)


def bm_l1_intended(ref, mov, flow, ts, r, return_margin=False):
    """The L1 level as block_matching.py:78-345 intends it (NOT what the compiled reference does, see bm_l1_compiled):
    E[v,u] = sum |ref - m| over the tile, m read at rint(flow) + (u, v) with ZERO outside the frame (:208-216), float64
    sums, first minimum in v-major order, flow <- rint(flow) + (u*, v*)."""
    ny, nx = flow.shape[:2]
    hm, wm = mov.shape
    f = np.rint(flow).astype(np.int64)
    reft = tile_view(ref.astype(F64), ts, ny, nx)
    n = 2 * r + 1
    E = np.empty((ny, nx, n, n), F64)
    ys = (np.arange(ny) * ts)[:, None, None, None] + np.arange(ts)[None, None, :, None]
    xs = (np.arange(nx) * ts)[None, :, None, None] + np.arange(ts)[None, None, None, :]
    movd = mov.astype(F64)
    for v in range(-r, r + 1):
        yr = ys + f[..., 1][:, :, None, None] + v
        for u in range(-r, r + 1):
            xr = xs + f[..., 0][:, :, None, None] + u
            inside = (yr >= 0) & (yr < hm) & (xr >= 0) & (xr < wm)
            m = np.where(inside, movd[np.clip(yr, 0, hm - 1), np.clip(xr, 0, wm - 1)], 0.0)
            E[:, :, v + r, u + r] = np.abs(reft - m).sum((-2, -1))
    Ef = E.reshape(ny, nx, n * n)
    idx = np.argmin(Ef, axis=-1)
    out = f.astype(F32)
    out[..., 0] += (idx % n - r).astype(F32)
    out[..., 1] += (idx // n - r).astype(F32)
    if return_margin:
        srt = np.sort(Ef, axis=-1)
        return out, (srt[..., 1] - srt[..., 0]) / np.maximum(np.abs(srt[..., 0]), 1e-12)
    return out


# --------------------------------------------------------------------------------------------------------------
# ICA — ICA.py:78-481
# --------------------------------------------------------------------------------------------------------------
def ica(ref, gx, gy, hess, mov, flow, ts, n_iter):
    """Inverse compositional Lucas-Kanade per tile; tile-size specific sampling rules:
    ts=8: clamped coordinates, float64 1/det (ICA.py:105-193); ts=16/32: zero fill (:195-369);
    ts=64: zero fill + the row off-by-one of the sliding window (:371-481)."""
    ny, nx = flow.shape[:2]
    hm, wm = mov.shape
    out = flow.astype(F32).copy()
    A00, A01, A10, A11 = hess[..., 0, 0], hess[..., 0, 1], hess[..., 1, 0], hess[..., 1, 1]
    det = A00 * A11 - A01 * A10
    if ts == 8:
        ok = ~(np.abs(det.astype(F64)) < 1e-10)
        with np.errstate(divide="ignore", invalid="ignore"):
            det_inv = 1.0 / det.astype(F64)
    else:
        ok = ~(np.abs(det) < F32(1e-10))
        with np.errstate(divide="ignore", invalid="ignore"):
            det_inv = F32(1.0) / det
    reft = tile_view(ref, ts, ny, nx)
    gxt = tile_view(gx, ts, ny, nx)
    gyt = tile_view(gy, ts, ny, nx)
    y0 = (np.arange(ny) * ts)[:, None, None, None] + np.arange(ts)[None, None, :, None]
    x0 = (np.arange(nx) * ts)[None, :, None, None] + np.arange(ts)[None, None, None, :]
    y0, x0 = np.broadcast_arrays(y0, x0)

    def S(yy, xx):
        inb = (yy >= 0) & (yy < hm) & (xx >= 0) & (xx < wm)
        return np.where(inb, mov[np.clip(yy, 0, hm - 1), np.clip(xx, 0, wm - 1)], F32(0))

    al = out.copy()
    for _ in range(n_iter):
        ix = np.trunc(al[..., 0]).astype(np.int64)[:, :, None, None]
        iy = np.trunc(al[..., 1]).astype(np.int64)[:, :, None, None]
        fx = (al[..., 0] - np.trunc(al[..., 0])).astype(F32)[:, :, None, None]
        fy = (al[..., 1] - np.trunc(al[..., 1])).astype(F32)[:, :, None, None]
        X, Y = x0 + ix, y0 + iy
        if ts == 8:
            Xf, Yf = np.clip(X, 0, wm - 1), np.clip(Y, 0, hm - 1)
            Xc, Yc = np.clip(Xf + 1, 0, wm - 1), np.clip(Yf + 1, 0, hm - 1)
            m00, m01, m10, m11 = mov[Yf, Xf], mov[Yf, Xc], mov[Yc, Xf], mov[Yc, Xc]
        elif ts == 64:
            i_in = (y0 - (np.arange(ny) * ts)[:, None, None, None]) % 4
            Ybase = Y - i_in                       # row of the thread's first pixel
            top_row = np.where(i_in == 0, Ybase, Ybase + i_in + 1)
            bot_row = Ybase + i_in + 2
            m00, m01 = S(top_row, X), S(top_row, X + 1)
            m10, m11 = S(bot_row, X), S(bot_row, X + 1)
        else:
            m00, m01, m10, m11 = S(Y, X), S(Y, X + 1), S(Y + 1, X), S(Y + 1, X + 1)
        top = m00 + (m01 - m00) * fx
        bot = m10 + (m11 - m10) * fx
        gt = (top + (bot - top) * fy) - reft
        B0 = (-gxt * gt).sum((-2, -1), dtype=F32)
        B1 = (-gyt * gt).sum((-2, -1), dtype=F32)
        with np.errstate(invalid="ignore", over="ignore"):
            d0 = det_inv * (A11 * B0 - A01 * B1)
            d1 = det_inv * (-A10 * B0 + A00 * B1)
        al[..., 0] = np.where(ok, (al[..., 0] + d0).astype(F32), al[..., 0])
        al[..., 1] = np.where(ok, (al[..., 1] + d1).astype(F32), al[..., 1])
    return al


# --------------------------------------------------------------------------------------------------------------
# Alignment driver — alignment.py
# --------------------------------------------------------------------------------------------------------------
def upscale_lvl(flow, npatchs, l, tile_sizes, factors, mode="nearest"):
    """alignment.py:150-172 (nearest mode; bilinear/bicubic are delegated to torch in the tests)."""
    new_ts, prev_ts, f = tile_sizes[l], tile_sizes[l + 1], factors[l + 1]
    rep = f // (new_ts // prev_ts)
    if mode != "nearest":
        import torch
        import torch.nn.functional as TF
        t = torch.from_numpy(flow).permute(2, 0, 1)[None]
        up = TF.interpolate(t, scale_factor=rep, mode=mode)[0].permute(1, 2, 0).numpy().copy()
    else:
        up = np.repeat(np.repeat(flow, rep, axis=0), rep, axis=1)
    up = (up * F32(f)).astype(F32)
    out = np.zeros((max(npatchs[0], up.shape[0]), max(npatchs[1], up.shape[1]), 2), F32)
    out[:up.shape[0], :up.shape[1]] = up
    return out


def pad_circular(img, ts):
    """alignment.py:26-37: pad bottom/right circularly to a multiple of the finest tile size."""
    h, w = img.shape
    ph = (ts - h % ts) * (h % ts != 0)
    pw = (ts - w % ts) * (w % ts != 0)
    return np.pad(img, ((0, ph), (0, pw)), mode="wrap")


def init_alignment(ref_grey, cfg):
    """alignment.py:20-72.  Returns dict with pyramid/gradients/hessians, coarse -> fine."""
    bm = cfg["block_matching"]["tuning"]
    factors, tss = bm["factors"], bm["tile_sizes"]
    pyr = build_gaussian_pyramid(pad_circular(ref_grey, bm["tile_size"]), factors)
    gxs, gys, hs = [], [], []
    for i, lvl in enumerate(pyr):
        ts = tss[len(factors) - i - 1]
        gx, gy, hh = init_ica(lvl, ts)
        gxs.append(gx), gys.append(gy), hs.append(hh)
    return dict(pyramid=pyr, gradx=gxs, grady=gys, hessian=hs)


def align(ref, mov_grey, cfg, trace=None):
    """alignment.py:84-147: coarse-to-fine {upscale, block matching, ICA}; returns flow [ny,nx,2] (dx,dy)."""
    bm = cfg["block_matching"]["tuning"]
    factors, tss, radii, metrics = bm["factors"], bm["tile_sizes"], bm["search_radii"], bm["metrics"]
    n_iter = cfg["ica"]["tuning"]["n_iter"]
    mpyr = build_gaussian_pyramid(mov_grey, factors)
    flow = None
    L = len(factors)
    for i in range(L):
        l = L - i - 1
        ts = tss[l]
        ref_lvl = ref["pyramid"][i]
        ny, nx = ref_lvl.shape[0] // ts, ref_lvl.shape[1] // ts
        if flow is None:
            flow = np.zeros((ny, nx, 2), F32)
        else:
            flow = upscale_lvl(flow, (ny, nx), l, tss, factors, bm.get("flow_upscale_mode", "nearest"))
        if trace is not None:
            trace["l%d_in" % l] = flow.copy()
        if metrics[l] == "L2":
            flow = bm_l2(ref_lvl, mpyr[i], flow, ts, radii[l])
        elif metrics[l] == "L1":
            flow = bm_l1_compiled(flow)
        else:
            raise ValueError("Unknown block matching metric")
        if trace is not None:
            trace["l%d_bm" % l] = flow.copy()
        flow = ica(ref_lvl, ref["gradx"][i], ref["grady"][i], ref["hessian"][i], mpyr[i], flow, ts, n_iter)
        if trace is not None:
            trace["l%d_ica" % l] = flow.copy()
    return flow


# --------------------------------------------------------------------------------------------------------------
# Kernel estimation (Alg. 5) — kernels.py, linalg.py:86-185
# --------------------------------------------------------------------------------------------------------------
def estimate_kernels(raw, cfg):
    """kernels.py:29-243.  Returns covs [H/2, W/2, 2, 2] f32."""
    mt = cfg["merging"]["tuning"]
    k_detail, k_denoise = float(mt["k_detail"]), float(mt["k_denoise"])
    D_th, D_tr = float(mt["D_th"]), float(mt["D_tr"])
    k_stretch, k_shrink = mt["k_stretch"], mt["k_shrink"]
    law = cfg["merging"]["selection_law"]
    g = decimate_to_grey(gat(raw, cfg["noise_model"]["alpha"], cfg["noise_model"]["beta"]))
    h, w = g.shape
    # two conv2d (kernels.py:97-112): horizontal [-.5,.5] / [.5,.5], then vertical [.5,.5] / [-.5,.5]
    t0 = F32(-0.5) * g[:, :-1] + F32(0.5) * g[:, 1:]
    t1 = F32(0.5) * g[:, :-1] + F32(0.5) * g[:, 1:]
    gx = F32(0.5) * t0[:-1] + F32(0.5) * t0[1:]
    gy = F32(-0.5) * t1[:-1] + F32(0.5) * t1[1:]
    T00 = np.zeros((h, w), F32)
    T01 = np.zeros((h, w), F32)
    T11 = np.zeros((h, w), F32)
    for i in range(2):
        for j in range(2):
            # grid point (y-1+i, x-1+j) must exist in the (h-1)x(w-1) gradient grid
            ys = slice(max(0, 1 - i), min(h, h - i))
            xs = slice(max(0, 1 - j), min(w, w - j))
            gys = slice(ys.start - 1 + i, ys.stop - 1 + i)
            gxs = slice(xs.start - 1 + j, xs.stop - 1 + j)
            a, b = gx[gys, gxs], gy[gys, gxs]
            T00[ys, xs] = fma32(a, a, T00[ys, xs])
            T01[ys, xs] = fma32(a, b, T01[ys, xs])
            T11[ys, xs] = fma32(b, b, T11[ys, xs])
    with np.errstate(all="ignore"):
        # eigenvalues, linalg.py:86-130 (a = 1 is an int: 4*a*c is float64)
        b_ = -(T00 + T11)
        c_ = fma32(T00, T11, -(T01 * T01))
        delta = np.maximum((b_ * b_).astype(F64) - 4.0 * c_.astype(F64), 0)
        sq = np.sqrt(delta)
        r1 = (-b_.astype(F64) + sq) / 2
        r2 = (-b_.astype(F64) - sq) / 2
        sw = np.abs(r1) >= np.abs(r2)
        l1 = np.where(sw, r1, r2).astype(F32)
        l2 = np.where(sw, r2, r1).astype(F32)
        # eigenvectors, linalg.py:132-179
        e1x = (T00 + T01 - l2).astype(F32)
        e1y = (T01 + T11 - l2).astype(F32)
        ident = (T01 == 0) & (T00 == T11)
        zx, zy = (e1x == 0), (e1y == 0)
        nrm = np.sqrt(fma32(e1x, e1x, e1y * e1y))
        nx_, ny_ = (e1x / nrm).astype(F32), (e1y / nrm).astype(F32)
        sign = np.copysign(F32(1), nx_)
        E1x, E1y = nx_, ny_
        E2x, E2y = (-ny_ * sign).astype(F32), np.abs(nx_)
        # e1y == 0 (and e1x != 0): e1 = (1, 0), e2 = (0, 1)
        E1x, E1y, E2x, E2y = (np.where(zy, F32(1), E1x), np.where(zy, F32(0), E1y),
                              np.where(zy, F32(0), E2x), np.where(zy, F32(1), E2y))
        # e1x == 0: e1 = (0, 1), e2 = (1, 0)   (tested first in the reference)
        E1x, E1y, E2x, E2y = (np.where(zx, F32(0), E1x), np.where(zx, F32(1), E1y),
                              np.where(zx, F32(1), E2x), np.where(zx, F32(0), E2y))
        E1x, E1y, E2x, E2y = (np.where(ident, F32(1), E1x), np.where(ident, F32(0), E1y),
                              np.where(ident, F32(0), E2x), np.where(ident, F32(1), E2y))
        # compute_k, kernels.py:194-243
        A = 1 + np.sqrt(((l1 - l2) / (l1 + l2)).astype(F32)).astype(F64)
        D = np.minimum(1, np.maximum(0, 1 - np.sqrt(l1).astype(F64) / D_tr + D_th))
        D = np.where(np.isnan(D), 0.0, D)  # clamp(NaN) = min(1, max(0, NaN)) = 0 with Numba's max/min
        if law == "hard_threshold":
            big = A > 1.95
            k1 = np.where(big, 1 / k_shrink, 1.0)
            k2 = np.where(big, float(k_stretch), 1.0)
        elif law == "linear":
            k1 = 1 + A / 2 * (1 / k_shrink - 1)
            k2 = 1 + A / 2 * (k_stretch - 1)
        else:
            raise ValueError("Unknown selection law: %s" % law)
        k1 = (k_detail * ((1 - D) * k1 + D * k_denoise)).astype(F32)
        k2 = (k_detail * ((1 - D) * k2 + D * k_denoise)).astype(F32)
        k1s, k2s = k1 * k1, k2 * k2
        covs = np.empty((h, w, 2, 2), F32)
        covs[..., 0, 0] = k1s * E1x * E1x + k2s * E2x * E2x
        covs[..., 0, 1] = covs[..., 1, 0] = k1s * E1x * E1y + k2s * E2x * E2y
        covs[..., 1, 1] = k1s * E1y * E1y + k2s * E2y * E2y
    return covs


# --------------------------------------------------------------------------------------------------------------
# Robustness (Alg. 6-9) — robustness.py
# --------------------------------------------------------------------------------------------------------------
def guide_image(raw, cfa, wb):
    """robustness.py:206-226: half-res RGB, white balance undone in float64, greens averaged."""
    h, w = raw.shape[0] // 2, raw.shape[1] // 2
    out = np.zeros((3, h, w), F32)
    g = np.zeros((h, w), F64)
    for i in range(2):
        for j in range(2):
            c = int(cfa[i][j])
            x = raw[i:2 * h:2, j:2 * w:2].astype(F64) / float(wb[c])
            if c == 1:
                g = g + x
            else:
                out[c] = x.astype(F32)
    out[1] = (g / 2).astype(F32)
    return out


def local_stats(guide):
    """robustness.py:268-294: 3x3 edge-replicated f32 sums, float64 finish."""
    c, h, w = guide.shape
    p = np.pad(guide, ((0, 0), (1, 1), (1, 1)), mode="edge")
    s1 = np.zeros_like(guide)
    s2 = np.zeros_like(guide)
    for i in range(3):
        for j in range(3):
            v = p[:, i:i + h, j:j + w]
            s1 += v
            s2 += v * v
    mean = s1.astype(F64) / 9
    var = s2.astype(F64) / 9 - mean * mean
    return mean.astype(F32), var.astype(F32)


def _dodgson(t):
    a = np.abs(t)
    return np.where(a <= 0.5, -2 * a * a + 1, np.where(a <= 1.5, a * a - 5 / 2 * a + 1.5, 0.0))


def upscale_warp_stats(LR, tile_size=None, flow=None):
    """robustness.py:358-418: x2 Dodgson biquadratic upsampling (+ tile-flow warp); +inf where the source
    position leaves the guide image."""
    c, lh, lw = LR.shape
    H, W = 2 * lh, 2 * lw
    y = np.arange(H)[:, None].astype(F64)
    x = np.arange(W)[None, :].astype(F64)
    if flow is None:
        fx = fy = 0.0
    else:
        ty = (np.arange(H) // tile_size)[:, None]
        tx = (np.arange(W) // tile_size)[None, :]
        fx = flow[ty, tx, 0].astype(F64)
        fy = flow[ty, tx, 1].astype(F64)
    ly = (y + fy + 0.5) / 2 - 0.5 + np.zeros((H, W))
    lx = (x + fx + 0.5) / 2 - 0.5 + np.zeros((H, W))
    inb = (ly >= 0) & (ly < lh) & (lx >= 0) & (lx < lw)
    cy = np.rint(np.where(inb, ly, 0)).astype(np.int64)
    cx = np.rint(np.where(inb, lx, 0)).astype(np.int64)
    buf = np.zeros((c, H, W), F32)
    wacc = np.zeros((H, W), F64)
    for i in (-1, 0, 1):
        y_ = np.clip(cy + i, 0, lh - 1)
        wy = _dodgson(y_ - ly)
        for j in (-1, 0, 1):
            x_ = np.clip(cx + j, 0, lw - 1)
            wgt = wy * _dodgson(x_ - lx)
            buf = (buf.astype(F64) + LR[:, y_, x_].astype(F64) * wgt).astype(F32)
            wacc = wacc + wgt
    with np.errstate(all="ignore"):
        out = (buf.astype(F64) / wacc).astype(F32)
    out[:, ~inb] = np.inf
    return out


def init_robustness(ref_raw, cfa, wb):
    """robustness.py:23-76."""
    m, s = local_stats(guide_image(ref_raw, cfa, wb))
    return upscale_warp_stats(m), upscale_warp_stats(s)


def compute_s(flow, Mt, s1, s2):
    """robustness.py:569-611."""
    ny, nx = flow.shape[:2]
    mx = np.full((ny, nx, 2), -np.inf, F32)
    mn = np.full((ny, nx, 2), np.inf, F32)
    p = np.pad(flow, ((1, 1), (1, 1), (0, 0)), mode="constant", constant_values=np.nan)
    for i in range(3):
        for j in range(3):
            v = p[i:i + ny, j:j + nx]
            mx = np.fmax(mx, v)
            mn = np.fmin(mn, v)
    d = mx - mn
    return np.where(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] > Mt * Mt, F32(s1), F32(s2)).astype(F32)


def local_min5(R):
    """robustness.py:669-687."""
    h, w = R.shape
    p = np.pad(R, 2, mode="edge")
    out = np.full_like(R, np.inf)
    for i in range(5):
        for j in range(5):
            out = np.minimum(out, p[i:i + h, j:j + w])
    return out


def compute_robustness(raw, ref_means, ref_stds, flow, cfa, wb, std_curve, diff_curve, cfg, return_R=False):
    """robustness.py:79-170 (+ :452-462, :504-533, :626-639)."""
    if not cfg["robustness"]["enabled"]:
        return np.ones_like(raw, F32)
    rt = cfg["robustness"]["tuning"]
    ts = cfg["block_matching"]["tuning"]["tile_size"]
    cm, _ = local_stats(guide_image(raw, cfa, wb))
    cm = upscale_warp_stats(cm, ts, flow)
    std_curve = np.asarray(std_curve, F64)
    diff_curve = np.asarray(diff_curve, F64)
    with np.errstate(all="ignore"):
        d_p = np.abs(ref_means - cm)                                   # f32
        fin = np.isfinite(ref_means)
        idx = np.rint(1000 * np.where(fin, ref_means, 0).astype(F64)).astype(np.int64)
        idx = np.clip(idx, 0, len(std_curve) - 1)
        sig_t, d_t = std_curve[idx], diff_curve[idx]
        sigma_sq = np.zeros(raw.shape, F64)
        d_sq = np.zeros(raw.shape, F64)
        for c in range(ref_means.shape[0]):
            sp = ref_stds[c].astype(F64)
            st2 = sig_t[c] * sig_t[c]
            sigma_sq = sigma_sq + np.where(st2 > sp, st2, sp)          # max(sigma_p_sq, sigma_t^2)
            dps = (d_p[c] * d_p[c]).astype(F32)                         # f32 * f32
            shrink = dps.astype(F64) / (dps.astype(F64) + d_t[c] * d_t[c])
            d_sq = d_sq + dps.astype(F64) * shrink * shrink
        sigma_sq, d_sq = sigma_sq.astype(F32), d_sq.astype(F32)
        S = compute_s(flow, rt["Mt"], rt["s1"], rt["s2"])
        H, W = raw.shape
        Sp = S[(np.arange(H) // ts)[:, None], (np.arange(W) // ts)[None, :]]
        e = np.exp((-d_sq / sigma_sq).astype(F32)).astype(F32)          # math.exp(float32) -> float32
        R = (Sp * e).astype(F32).astype(F64) - rt["t"]
        R = np.where(np.isnan(R), 0.0, np.minimum(1, np.maximum(0, R))).astype(F32)
    r = local_min5(R)
    return (r, R) if return_R else r


# --------------------------------------------------------------------------------------------------------------
# Merge (Alg. 4, Alg. 11) — merge.py, linalg.py:37-84,189-200
# --------------------------------------------------------------------------------------------------------------
def _cfa_channel(cfa, i, j):
    cfa = np.asarray(cfa)
    return cfa[i % 2, j % 2]


def accumulate(raw, flow, covs, r, num, den, cfa, scale, tile_size, iso_kernel=False):
    """merge.py:290-434 — in place on num/den [Hs,Ws,3] f32.  All position/weight math in float64."""
    hr_h, hr_w = num.shape[:2]
    lr_h, lr_w = raw.shape
    hr_i, hr_j = np.meshgrid(np.arange(hr_h), np.arange(hr_w), indexing="ij")
    lr_x = (hr_j + 0.5) / scale
    lr_y = (hr_i + 0.5) / scale
    px = (lr_x // tile_size).astype(np.int64)
    py = (lr_y // tile_size).astype(np.int64)
    flowx = flow[py, px, 0].astype(F64)
    flowy = flow[py, px, 1].astype(F64)
    local_r = r[np.minimum(lr_y.astype(np.int64), lr_h - 1), np.minimum(lr_x.astype(np.int64), lr_w - 1)].astype(F64)
    mx, my = lr_x + flowx, lr_y + flowy
    inb = (mx >= 0) & (mx < lr_w) & (my >= 0) & (my < lr_h)
    mxs, mys = np.where(inb, mx, 0.0), np.where(inb, my, 0.0)
    with np.errstate(all="ignore"):
        if not iso_kernel:
            kj, ki = mxs / 2 - 0.5, mys / 2 - 0.5
            frx, fry = kj - np.trunc(kj), ki - np.trunc(ki)
            fx0 = np.maximum(np.trunc(kj).astype(np.int64), 0)
            fy0 = np.maximum(np.trunc(ki).astype(np.int64), 0)
            cx1 = np.minimum(fx0 + 1, covs.shape[1] - 1)
            cy1 = np.minimum(fy0 + 1, covs.shape[0] - 1)

            def interp(a, b):
                tr = covs[fy0, fx0, a, b].astype(F64)
                tl = covs[fy0, cx1, a, b].astype(F64)
                br = covs[cy1, fx0, a, b].astype(F64)
                bl = covs[cy1, cx1, a, b].astype(F64)
                top = tr + frx * (tl - tr)
                bot = br + frx * (bl - br)
                return top + fry * (bot - top)
            cxx, cxy, cyy = interp(0, 0), interp(0, 1), interp(1, 1)
            inv_det = 1.0 / (cxx * cyy - cxy * cxy)
            ixx, ixy, iyy = inv_det * cyy, -inv_det * cxy, inv_det * cxx
        cj = np.trunc(mxs).astype(np.int64)
        ci = np.trunc(mys).astype(np.int64)
        mj, mi = mxs - 0.5, mys - 0.5
        val = np.zeros((hr_h, hr_w, 3), F32)
        acc = np.zeros((hr_h, hr_w, 3), F32)
        cfa = np.asarray(cfa)
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                j, i = cj + dj, ci + di
                ok = inb & (j >= 0) & (j < lr_w) & (i >= 0) & (i < lr_h)
                jc, ic = np.clip(j, 0, lr_w - 1), np.clip(i, 0, lr_h - 1)
                ch = cfa[ic % 2, jc % 2]
                c = raw[ic, jc].astype(F64)
                dx, dy = j - mj, i - mi
                if iso_kernel:
                    z = 2 * (dx * dx + dy * dy)
                else:
                    z = ixx * dx * dx + 2 * ixy * dx * dy + iyy * dy * dy
                z = np.where(z > 0, z, 0.0)                      # max(0, z): NaN -> 0
                w = np.exp(-0.5 * z)
                wr = w * local_r
                for k in range(3):
                    m = ok & (ch == k)
                    val[..., k] = np.where(m, (val[..., k].astype(F64) + wr * c).astype(F32), val[..., k])
                    acc[..., k] = np.where(m, (acc[..., k].astype(F64) + wr).astype(F32), acc[..., k])
    num += val
    den += acc


def accumulate_ref(raw, covs, num, den, cfa, scale, iso_kernel=False, acc_rob=None, max_frame_count=0,
                   rad_max=0, max_multiplier=0.0):
    """merge.py:82-233 (+ linalg.py:37-84,189-200, utils_image.py:311-325) — in place on num/den."""
    hr_h, hr_w = num.shape[:2]
    lr_h, lr_w = raw.shape
    oy, ox = np.meshgrid(np.arange(hr_h), np.arange(hr_w), indexing="ij")
    py = (oy / scale).astype(F32)
    px = (ox / scale).astype(F32)
    with np.errstate(all="ignore"):
        if not iso_kernel:
            gy = ((py.astype(F64) - 0.5) / 2).astype(F32)
            gx = ((px.astype(F64) - 0.5) / 2).astype(F32)
            fx0 = np.maximum(np.floor(gx), 0).astype(np.int64)
            fy0 = np.maximum(np.floor(gy), 0).astype(np.int64)
            cx1 = np.minimum(fx0 + 1, covs.shape[1] - 1)
            cy1 = np.minimum(fy0 + 1, covs.shape[0] - 1)
            rx = (gx - np.trunc(gx)).astype(F32).astype(F64)
            ry = (gy - np.trunc(gy)).astype(F32).astype(F64)
            ic = np.empty((hr_h, hr_w, 2, 2), F32)
            for a in range(2):
                for b in range(2):
                    ic[..., a, b] = (covs[fy0, fx0, a, b].astype(F64) * (1 - rx) * (1 - ry) +
                                     covs[fy0, cx1, a, b].astype(F64) * rx * (1 - ry) +
                                     covs[cy1, fx0, a, b].astype(F64) * (1 - rx) * ry +
                                     covs[cy1, cx1, a, b].astype(F64) * rx * ry).astype(F32)
            det = fma32(ic[..., 0, 0], ic[..., 1, 1], -(ic[..., 0, 1] * ic[..., 1, 0]))   # f32, contracted by NVVM
            good = np.abs(det.astype(F64)) > EPSILON_DIV
            det_i = 1 / det.astype(F64)
            i00 = np.where(good, (ic[..., 1, 1] * det_i).astype(F32), F32(1))
            i01 = np.where(good, (-ic[..., 0, 1] * det_i).astype(F32), F32(0))
            i10 = np.where(good, (-ic[..., 1, 0] * det_i).astype(F32), F32(0))
            i11 = np.where(good, (ic[..., 0, 0] * det_i).astype(F32), F32(1))
        denoise = acc_rob is not None
        if denoise:
            la = acc_rob[np.minimum(np.rint(py).astype(np.int64), acc_rob.shape[0] - 1),
                         np.minimum(np.rint(px).astype(np.int64), acc_rob.shape[1] - 1)]
            few = la <= max_frame_count
            power = np.where(few, float(max_multiplier), 1.0)
            rad = np.where(few, rad_max, 1)
            R = max(int(rad_max), 1)
        else:
            power, rad, R = 1.0, None, 1
        cx = np.rint(px).astype(np.int64)
        cy = np.rint(py).astype(np.int64)
        val = np.zeros((hr_h, hr_w, 3), F32)
        acc = np.zeros((hr_h, hr_w, 3), F32)
        cfa = np.asarray(cfa)
        for i in range(-R, R + 1):
            for j in range(-R, R + 1):
                xx, yy = cx + j, cy + i
                ok = (xx >= 0) & (xx < lr_w) & (yy >= 0) & (yy < lr_h)
                if rad is not None:
                    ok &= (abs(i) <= rad) & (abs(j) <= rad)
                xc, yc = np.clip(xx, 0, lr_w - 1), np.clip(yy, 0, lr_h - 1)
                ch = cfa[yc % 2, xc % 2]
                c = raw[yc, xc].astype(F64)
                dx = xx - px.astype(F64)
                dy = yy - py.astype(F64)
                if iso_kernel:
                    y = 2 * (dx * dx + dy * dy)
                else:
                    y = i00.astype(F64) * dx * dx + dx * dy * (i01 + i10).astype(F64) + i11.astype(F64) * dy * dy
                y = np.where(y > 0, y, 0.0)
                y = y / power
                w = np.exp(-0.5 * y)
                for k in range(3):
                    m = ok & (ch == k)
                    val[..., k] = np.where(m, (val[..., k].astype(F64) + c * w).astype(F32), val[..., k])
                    acc[..., k] = np.where(m, (acc[..., k].astype(F64) + w).astype(F32), acc[..., k])
    if denoise:
        over = (la < max_frame_count)[..., None]
        num[...] = np.where(over, val, num + val)
        den[...] = np.where(over, acc, den + acc)
    else:
        num += val
        den += acc


def divide(num, den):
    """utils.py:84-90."""
    with np.errstate(all="ignore"):
        return (num / den).astype(F32)


# --------------------------------------------------------------------------------------------------------------
# Alg. 1 — super_resolution.py:41-200
# --------------------------------------------------------------------------------------------------------------
def main(ref_img, comp_imgs, cfg, trace=None):
    """Whole device pipeline; cfg is a plain nested dict with the reference's keys.  Returns (out, debug)."""
    scale = cfg["scale"]
    ts = cfg["block_matching"]["tuning"]["tile_size"]
    cfa, wb = cfg["exif"]["cfa_pattern"], cfg["exif"]["white_balance"]
    iso = cfg["merging"]["kernel"] == "iso"
    stdc, diffc = cfg["noise_model"]["std_curve"], cfg["noise_model"]["diff_curve"]
    ref = init_alignment(grey_fft(ref_img), cfg)
    r_on = cfg["robustness"]["enabled"]
    if r_on:
        rm, rs = init_robustness(ref_img, cfa, wb)
    H, W = ref_img.shape
    hs, ws = round(scale * H), round(scale * W)
    num = np.zeros((hs, ws, 3), F32)
    den = np.zeros((hs, ws, 3), F32)
    acc_rob = np.zeros((H, W), F64)
    dbg = {"flow": [], "robustness": []}
    for img in comp_imgs:
        flow = align(ref, grey_fft(img), cfg)
        r = compute_robustness(img, rm, rs, flow, cfa, wb, stdc, diffc, cfg) if r_on else np.ones_like(img, F32)
        acc_rob += r
        covs = estimate_kernels(img, cfg)
        accumulate(img, flow, covs, r, num, den, cfa, scale, ts, iso)
        dbg["flow"].append(flow), dbg["robustness"].append(r)
    covs = estimate_kernels(ref_img, cfg)
    ard = cfg.get("accumulated_robustness_denoiser", {})
    if ard.get("enabled", False):
        m = ard["merge"]
        accumulate_ref(ref_img, covs, num, den, cfa, scale, iso, acc_rob, m["max_frame_count"], m["rad_max"],
                       m["max_multiplier"])
    else:
        accumulate_ref(ref_img, covs, num, den, cfa, scale, iso)
    dbg["accumulated robustness"] = acc_rob
    dbg["num"], dbg["den"] = num.copy(), den.copy()
    return divide(num, den), dbg


# --------------------------------------------------------------------------------------------------------------
# Output side — raw2rgb.py:139-250, run_handheld.py:132-150, utils_image.py:174-315
#
# Parity status of this block.  raw2rgb.postprocess is pinned to the reference's own function (tests/golden/
# post_cases.npz, generator tests/golden/make_golden_post.py, which runs /root/reference's postprocess() with
# skimage.filters.unsharp_mask — scikit-image is not installed here — restated on top of the real
# scipy.ndimage.gaussian_filter it calls).  frame_count_denoising_median is pinned to the reference's kernel run under
# NUMBA_ENABLE_CUDASIM (same fixture).  frame_count_denoising_gauss is PARITY UNPINNED: upstream it cannot execute
# (range() over a float, utils_image.py:210-215), neither compiled nor simulated; the one repair (t = ceil(3*sigma)) is ours.
# --------------------------------------------------------------------------------------------------------------
def gaussian_taps(sigma, radius):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) (scipy 1.18): float64, normalised, symmetric."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def gaussian_filter_reflect(img, sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter(img, sigma, mode='reflect', truncate=4) of a 2-D float32 image: correlate1d along
    axis 0, then along axis 1; float64 accumulation, each pass rounded to the input dtype (scipy keeps it)."""
    radius = int(truncate * float(sigma) + 0.5)
    w = gaussian_taps(sigma, radius)
    out = np.asarray(img)
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (radius, radius)
        p = np.pad(out.astype(F64), pad, mode="symmetric")      # numpy 'symmetric' == scipy 'reflect' (d c b a | a b c d | d c b a)
        acc = np.zeros(out.shape, F64)
        n = out.shape[axis]
        for k in range(2 * radius + 1):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(k, k + n)
            acc += w[k] * p[tuple(sl)]
        out = acc.astype(img.dtype)
    return out


def unsharp_mask(img, radius, amount):
    """skimage.filters.unsharp_mask(img, radius, amount, channel_axis=2, preserve_range=True) (scikit-image 0.2x,
    filters/_unsharp_mask.py): per channel  image + (image - gaussian(image, sigma=radius, mode='reflect')) * amount,
    float32 in -> float32 out, no clipping."""
    img = np.asarray(img, F32)
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        blurred = gaussian_filter_reflect(img[..., c], radius)
        out[..., c] = img[..., c] + (img[..., c] - blurred) * F32(amount)
    return out


def color_matrix(xyz2cam):
    """raw2rgb.py:118-136 get_color_matrix -> cam2rgb (:223-224)."""
    rgb2xyz = np.array([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]])
    xyz2cam = np.asarray(xyz2cam)
    rgb2cam = rgb2xyz if np.linalg.norm(xyz2cam) == 0 else xyz2cam @ rgb2xyz
    rgb2cam = (rgb2cam / rgb2cam.sum(axis=-1, keepdims=True)).astype(F32)
    return np.linalg.inv(rgb2cam)


def postprocess(img, do_color_correction=True, do_tonemapping=False, do_gamma=True, sharpening=None, do_devignette=False,
                xyz2cam=None):
    """raw2rgb.py:212-250 (the `img is not None` branch) on a float32 [H,W,3] image."""
    img = np.asarray(img, F32)
    if do_color_correction:
        cam2rgb = color_matrix(xyz2cam)
        img = np.clip(np.einsum("ij,hwj->hwi", cam2rgb, img).astype(F32), 0.0, 1.0)         # apply_ccm :139-146
    if sharpening is not None and sharpening.get("enabled", False):
        img = unsharp_mask(img, sharpening.get("radius", 3), sharpening.get("amount", 0.5))
    if do_devignette:                                                                          # :203-210 (float64 from here on)
        h, w, _ = img.shape
        vf = np.abs(np.linspace(-h / w * np.pi / 2, h / w * np.pi / 2, h))
        vf = np.outer(vf, np.abs(np.linspace(-np.pi / 2, np.pi / 2, w)))
        img = (2 - np.cos(vf) ** 4)[:, :, None] * img
    if do_tonemapping:
        raise NotImplementedError("apply_smoothstep (OpenCV MergeMertens, raw2rgb.py:153-170) is outside the restated path")
    img = np.clip(img, 0.0, 1.0)
    if do_gamma:
        img = np.clip(img, 0.0, 1.0) ** (1.0 / 2.2)                                            # gamma_compression :143-146
    return np.clip(img, 0.0, 1.0)


def img_as_ubyte(img, top=255):
    """run_handheld.py:132-133,150: nan_to_num, clip, skimage.img_as_ubyte of a float32 image = rint(x * 255) in float32."""
    x = np.clip(np.nan_to_num(np.asarray(img, F32)), 0, 1)
    return np.rint(x * F32(top)).astype(np.uint8 if top == 255 else np.uint16)


def _grey_index(v, scale, n):
    return min(max(int(round((v - 0.5) / (2 * scale))), 0), n - 1)                            # utils_image.py:204-205 (Python round: half to even)


def frame_count_denoising_gauss(image, r_acc, scale, sigma_max, max_frame_count):
    """utils_image.py:192-231 with t = ceil(3 sigma) (upstream iterates range() over the float 3*sigma and cannot run)."""
    image = np.asarray(image, F32)
    Hs, Ws, _ = image.shape
    out = np.empty_like(image)
    for y in range(Hs):
        for x in range(Ws):
            r = min(r_acc[_grey_index(y, scale, r_acc.shape[0]), _grey_index(x, scale, r_acc.shape[1])], max_frame_count)
            sigma = sigma_max * (max_frame_count - r) / max_frame_count
            t = int(math.ceil(3 * sigma))
            if t <= 0:
                out[y, x] = image[y, x]
                continue
            num, den = np.zeros(3, F64), 0.0
            for i in range(-t, t + 1):
                for j in range(-t, t + 1):
                    if 0 <= y + i < Hs and 0 <= x + j < Ws:
                        w = math.exp(-(j * j + i * i) / (2 * sigma * sigma))
                        num += w * image[y + i, x + j].astype(F64)
                        den += w
            out[y, x] = (num / den).astype(F32)
    return out


def frame_count_denoising_median(image, r_acc, scale, radius_max, max_frame_count):
    """utils_image.py:251-315: window radius from the accumulated robustness, literal bubble sort, buffer[k // 2]."""
    image = np.asarray(image, F32)
    Hs, Ws, C = image.shape
    out = np.empty_like(image)
    for y in range(Hs):
        for x in range(Ws):
            r = min(r_acc[_grey_index(y, scale, r_acc.shape[0]), _grey_index(x, scale, r_acc.shape[1])], max_frame_count)
            radius = min(14, int(round(radius_max * (max_frame_count - r) / max_frame_count)))
            for c in range(C):
                buf = [image[y + i, x + j, c] for i in range(-radius, radius + 1) for j in range(-radius, radius + 1)
                       if 0 <= y + i < Hs and 0 <= x + j < Ws]
                k = len(buf)
                for i in range(k - 1):
                    for j in range(k - i - 1):
                        if buf[j] > buf[j + 1]:
                            buf[j], buf[j + 1] = buf[j + 1], buf[j]
                out[y, x, c] = buf[k // 2]
    return out
