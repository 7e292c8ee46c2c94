"""Element-wise helpers and constants — mirrors handheld_super_resolution/utils.py of the reference
(divide :62-90, add :93-120, dtype constants :16-23)."""
import time

import numpy as np
import torch

from . import _lib

DEFAULT_NUMPY_FLOAT_TYPE = np.float32
DEFAULT_TORCH_FLOAT_TYPE = torch.float32
DEFAULT_TORCH_COMPLEX_TYPE = torch.complex64
EPSILON_DIV = 1e-10
DEFAULT_THREADS = 16


def divide(num, den):
    """num <- num / den in place (utils.py:62-90); 0/0 gives NaN like the reference."""
    assert num.shape == den.shape
    _lib.call("hhsr_divide", _lib.ptr(num), _lib.ptr(den), num.numel(), _lib.stream())


def add(A, B):
    """A (float64) += B (float32) in place (utils.py:93-120)."""
    assert A.shape == B.shape
    assert A.dtype == torch.float64 and B.dtype == torch.float32
    _lib.call("hhsr_add_f64_f32", _lib.ptr(A), _lib.ptr(B), A.numel(), _lib.stream())


def add_many(A, Bs):
    """A (float64) += B_0 + B_1 + ... (float32, list order) in one pass over A — the per-frame `add(accumulated_r, r)`
    of super_resolution.py:159 deferred to the end of the frame loop (same float64 sums, 1/K of the traffic on A)."""
    import ctypes as C
    assert A.dtype == torch.float64 and all(b.dtype == torch.float32 and b.shape == A.shape for b in Bs)
    if not Bs:
        return
    arr = (C.c_void_p * len(Bs))(*[b.data_ptr() for b in Bs])
    _lib.call("hhsr_add_many_f64_f32", _lib.ptr(A), arr, len(Bs), A.numel(), _lib.stream())


def getTime(currentTime, labelName, printTime=True, spaceSize=50):
    """utils.py:26-30."""
    if printTime:
        print(labelName, " " * (spaceSize - len(labelName)), ": ",
              round((time.perf_counter() - currentTime) * 1000, 2), "milliseconds")
    return time.perf_counter()


def timer(func, enabled, start_s=None, end_s=None, spaceSize=50):
    """utils.py:128-146: wall-clock around device-wide synchronisations, only when `enabled`."""
    if not enabled:
        return func

    def wrapper(*args, **kwargs):
        torch.cuda.synchronize()
        if start_s is not None:
            print(start_s)
        t = time.perf_counter()
        out = func(*args, **kwargs)
        torch.cuda.synchronize()
        if end_s is not None:
            getTime(t, end_s, True, spaceSize)
        return out
    return wrapper


def round_iso(iso):
    """utils.py:122-125."""
    import math
    return int(100 * (2 ** round(math.log2(iso / 100))))
