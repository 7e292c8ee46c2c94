"""CPU check of the grey-image FFT kernels' arithmetic: csrc/fft_core.cuh (the per-thread phases the CUDA kernels of
csrc/greyfft.cu are made of) is compiled with g++ into an emulation that runs the threads of each CTA one after the
other (tests/native/greyfft_emul.cpp) and compared with numpy.fft and with the oracle's grey image."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("greyfft") / "greyfft_emul.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                    "-I", os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "greyfft_emul.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def test_factorization(emul):
    """Few large radices; even ones first, the odd one last (no padding then); sizes without an odd factor are padded."""
    radix, pad = (C.c_int * 16)(), C.c_int()
    for n, want, want_pad in [(4000, [16, 10, 25], 0), (3000, [20, 10, 15], 0), (8192, [32, 16, 16], 1), (6144, [24, 16, 16], 1),
                              (1400, [20, 10, 7], 0), (2, [2], 1), (3024, None, None), (4032, None, None), (6240, [24, 20, 13], 0),
                              (5472, [24, 12, 19], 0), (4624, [16, 17, 17], 0), (66, [6, 11], 0)]:
        cnt = emul.emul_factorize(n, radix, C.byref(pad))
        got = list(radix[:cnt])
        assert int(np.prod(got)) == n
        if want is not None:
            assert sorted(got) == sorted(want), (n, got)
        if want_pad is not None:
            assert pad.value == want_pad, (n, got)
        evens = [r for r in got if r % 2 == 0]
        assert got[:len(evens)] == sorted(evens, reverse=True) and all(r % 2 for r in got[len(evens):]), got
        assert pad.value == (0 if got[-1] % 2 else 1)
    # common sensor formats: which run the native passes ...
    for n in (4000, 3000, 4032, 3024, 6000, 4608, 3456, 8192, 6144, 5472, 3648, 4624, 6240, 4160):
        assert emul.emul_factorize(n, radix, C.byref(pad)) == 3, n
    assert emul.emul_factorize(9504, radix, C.byref(pad)) == 4 and emul.emul_factorize(6336, radix, C.byref(pad)) == 3
    # ... and which keep the cuFFT route: a prime factor above 19 (31 | 3472, 43 | 8256, 37 | 740)
    for n in (1, 3472, 8256, 740, 23, 62):
        assert emul.emul_factorize(n, radix, C.byref(pad)) == -1, n


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 6, 9, 10, 12, 15, 20, 21, 24, 25, 30, 60, 100, 125, 250, 256, 14, 49, 210, 441, 11, 13, 17, 19, 22, 26, 121, 143, 187, 361, 3000, 5472, 3648, 4624, 6240, 4160, 9504, 4000, 6144, 8192, 1400, 2560, 3024, 4032, 65536])
def test_transform_against_numpy(emul, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for inverse in (0, 1):
        for nt in (1, 7, 64):                # the result must not depend on how butterflies are dealt to threads
            d = x.copy()
            assert emul.emul_fft(d.ctypes.data_as(C.c_void_p), n, inverse, nt) == 0
            ref = np.fft.ifft(x.astype(np.complex128)) * n if inverse else np.fft.fft(x.astype(np.complex128))
            assert np.abs(d - ref).max() / np.abs(ref).max() < 4e-7, (n, inverse, nt)


def test_kept_columns(emul):
    import hhsr_oracle as O
    for W in (8, 24, 56, 64, 72, 128, 168, 1000, 4000, 8192):
        keep, keep_neg = O.grey_band_weights(W)
        wx = (keep | keep_neg)[: W // 2 + 1]
        assert emul.emul_kept_columns(W) == int(np.nonzero(wx)[0].max()) + 1, W


@pytest.mark.parametrize("H,W,CW", [(16, 24, 8), (48, 72, 4), (120, 168, 8), (96, 128, 2), (100, 56, 1), (30, 50, 8), (350, 360, 8), (66, 104, 8), (34, 38, 4), (208, 312, 8),
                                    (750, 1000, 8)])
def test_grey_image_against_oracle(emul, H, W, CW):
    """rows forward (row pairs, pruned store) -> columns (forward, mask, inverse) -> rows inverse == the reference's
    fft2 / fftshift / masked fills / ifftshift / ifft2 / .real (oracle.grey_fft, float64 numpy)."""
    import hhsr_oracle as O
    rng = np.random.default_rng(H * W)
    img = rng.random((H, W)).astype(np.float32)
    out = np.full_like(img, np.nan)
    assert emul.emul_grey(img.ctypes.data_as(C.c_void_p), H, W, out.ctypes.data_as(C.c_void_p), CW, 32, 64) == 0
    assert np.abs(out - O.grey_fft(img)).max() < 1e-6
