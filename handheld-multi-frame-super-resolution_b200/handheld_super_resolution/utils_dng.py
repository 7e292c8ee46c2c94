"""RAW normalisation — the device-side counterpart of the loop that closes load_dng_burst() in the reference
(handheld_super_resolution/utils_dng.py:146-160): black/white level and white-balance normalisation of the sensor
counts.  DNG decoding itself (rawpy / exifread) is host file I/O and stays out of scope;
what is kept is the arithmetic, so that a burst can cross PCIe as uint16 (2 B per pixel instead of 4) and be
normalised on the GPU — bit-identical to the reference's NumPy float32 chain."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class RawNormalization:
    """Per-CFA-position constants of utils_dng.py:151-157: black level, white - black, white-balance gain
    wb[c] / wb[1], each rounded to float32 exactly where NumPy rounds the Python scalars."""

    def __init__(self, cfa_pattern, black_levels, white_level, white_balance):
        cfa = np.asarray(cfa_pattern).astype(np.int64).reshape(2, 2)
        black = [float(b) for b in np.asarray(black_levels).reshape(-1)]
        wb = [float(w) for w in np.asarray(white_balance, dtype=np.float64).reshape(-1)]
        self.black, self.den, self.gain = [], [], []
        for i in range(2):
            for j in range(2):
                c = int(cfa[i, j])
                self.black.append(np.float32(black[c]))
                self.den.append(np.float32(float(white_level) - black[c]))
                self.gain.append(np.float32(wb[c] / wb[1]))

    @classmethod
    def from_config(cls, config):
        ex = config.exif
        if ex.get("black_levels", None) is None or ex.get("white_level", None) is None:
            raise ValueError("uint16 RAW frames need config.exif.black_levels and config.exif.white_level")
        return cls(ex.cfa_pattern, ex.black_levels, ex.white_level, ex.white_balance)

    def _arr(self, v):
        return (C.c_float * 4)(*[float(x) for x in v])

    def apply(self, raw_u16, out=None):
        """raw_u16: CUDA uint16 tensor [H, W] -> float32 [H, W] (one launch on the current stream)."""
        assert raw_u16.is_cuda and raw_u16.dtype == torch.uint16 and raw_u16.is_contiguous()
        H, W = raw_u16.shape
        if out is None:
            out = torch.empty((H, W), dtype=torch.float32, device=raw_u16.device)
        _lib.call("hhsr_normalize_raw_u16", _lib.ptr(raw_u16), H, W, self._arr(self.black), self._arr(self.den),
                  self._arr(self.gain), _lib.ptr(out), _lib.stream())
        return out

    def apply_numpy(self, raw):
        """The reference's host arithmetic (utils_dng.py:146-160) on a [..., H, W] integer array — what the device
        kernel must reproduce bit for bit; used by the tests and by process() when frames are already float."""
        x = np.asarray(raw).astype(np.float32)
        for i in range(2):
            for j in range(2):
                p = 2 * i + j
                x[..., i::2, j::2] = (x[..., i::2, j::2] - self.black[p]) / self.den[p]
                x[..., i::2, j::2] *= self.gain[p]
        return x
