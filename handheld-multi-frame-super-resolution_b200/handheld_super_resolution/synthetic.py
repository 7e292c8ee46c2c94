"""Synthetic Bayer bursts (SURVEY.md section 8d "Synthetic input").

The reference ships no test bursts (test_burst/readme.txt:1-2), so every parity test and benchmark of this
repository runs on bursts made here: a band-limited random RGB scene rendered at 2x, one global sub-pixel
translation per frame, 2x box decimation, RGGB mosaic and heteroscedastic Gaussian noise N(0, alpha*I + beta)
(the reference's own noise model, README.md:159-160 of the reference, ISO 100 values).

Runs on any torch device: CPU for the committed golden fixtures, CUDA for the 12 MP / 50 MP benchmark bursts.
"""
import numpy as np
import torch
import torch.nn.functional as F

ALPHA_ISO100 = 1.80710882e-4
BETA_ISO100 = 3.1937599182128e-6
CFA_RGGB = [[0, 1], [1, 2]]
WHITE_BALANCE = [2.0, 1.0, 1.5, 0.0]


def synth_burst(n, H, W, seed=0, max_shift=3.0, device="cpu", alpha=ALPHA_ISO100, beta=BETA_ISO100,
                quantize_bits=None, as_numpy=True):
    """Return (burst [n,H,W] float32 in [0,1], shifts [(dy,dx)...]).  frame[y,x] = scene[y-dy, x-dx] so the
    reference's flow convention gives flow == +(dx, dy).  Frame 0 is unshifted (the reference frame)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    up, pad = 2, 16
    hh, ww = H * up + 2 * pad, W * up + 2 * pad
    coarse = torch.rand((1, 3, hh // 4 + 2, ww // 4 + 2), device=dev, generator=g)
    coarse = F.interpolate(coarse, size=(hh, ww), mode="bicubic", align_corners=False)
    fine = torch.rand((1, 3, hh, ww), device=dev, generator=g)
    k = torch.tensor([1, 4, 6, 4, 1.0], device=dev) / 16
    fine = F.conv2d(fine, k.view(1, 1, 5, 1).repeat(3, 1, 1, 1), groups=3, padding=(2, 0))
    fine = F.conv2d(fine, k.view(1, 1, 1, 5).repeat(3, 1, 1, 1), groups=3, padding=(0, 2))
    scene = (0.6 * coarse + 0.4 * fine).clamp(0, 1) * 0.8 + 0.05
    del coarse, fine
    rng = np.random.default_rng(seed)
    ys = torch.arange(H * up, device=dev, dtype=torch.float32)
    xs = torch.arange(W * up, device=dev, dtype=torch.float32)
    frames, shifts = [], []
    for i in range(n):
        dy, dx = (0.0, 0.0) if i == 0 else rng.uniform(-max_shift, max_shift, 2)
        shifts.append((float(dy), float(dx)))
        gy = (ys + pad - dy * up + 0.5) / hh * 2 - 1
        gx = (xs + pad - dx * up + 0.5) / ww * 2 - 1
        grid = torch.stack(torch.broadcast_tensors(gx[None, :], gy[:, None]), -1)[None]
        lr = F.avg_pool2d(F.grid_sample(scene, grid, mode="bilinear", align_corners=False), up)[0]
        del grid
        bay = torch.empty((H, W), device=dev)
        bay[0::2, 0::2] = lr[0, 0::2, 0::2]
        bay[0::2, 1::2] = lr[1, 0::2, 1::2]
        bay[1::2, 0::2] = lr[1, 1::2, 0::2]
        bay[1::2, 1::2] = lr[2, 1::2, 1::2]
        noise = torch.randn((H, W), device=dev, generator=g)
        bay = (bay + torch.sqrt(alpha * bay + beta) * noise).clamp(0, 1)
        if quantize_bits:
            q = float(2 ** quantize_bits - 1)
            bay = torch.round(bay * q) / q
        frames.append(bay)
    burst = torch.stack(frames)
    if as_numpy:
        return burst.cpu().numpy().astype(np.float32), shifts
    return burst, shifts
