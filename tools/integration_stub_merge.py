"""The ctypes stub of INTEGRATION.md section 2, verbatim: what a maintainer of the reference would paste into the reference's
merge.py to replace the `accumulate[blockspergrid, threadsperblock](...)` launch (merge.py:284-287) by libhhsr.so.
tests/test_gpu_stage_swap.py monkey-patches it into the UNMODIFIED reference and compares whole-pipeline outputs."""
import ctypes
import os

import numpy as np

_hhsr = ctypes.CDLL(os.environ.get("HHSR_LIB", "libhhsr.so"))
_hhsr.hhsr_merge_accumulate.restype = ctypes.c_int
_hhsr.hhsr_last_error_string.restype = ctypes.c_char_p
_P, _I, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_hhsr.hhsr_merge_accumulate.argtypes = [_P, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _D,
                                        ctypes.POINTER(_I), _I, _P]


def _ptr(a):                      # Numba DeviceNDArray or torch tensor
    return _P(a.__cuda_array_interface__["data"][0])


def merge(comp_img, alignments, covs, r, num, den, cfa_pattern, config):
    H, W = comp_img.shape
    ny, nx, _ = alignments.shape
    cfa = (ctypes.c_int * 4)(*np.asarray(cfa_pattern.copy_to_host()).astype(int).ravel())
    rc = _hhsr.hhsr_merge_accumulate(_ptr(comp_img), H, W, _ptr(alignments), ny, nx,
                                     config.block_matching.tuning.tile_size, _ptr(covs), _ptr(r), _ptr(num), _ptr(den),
                                     num.shape[0], num.shape[1], float(config.scale), cfa,
                                     int(config.merging.kernel == "iso"), None)   # NULL = Numba's default stream
    if rc:
        raise RuntimeError(_hhsr.hhsr_last_error_string().decode())
