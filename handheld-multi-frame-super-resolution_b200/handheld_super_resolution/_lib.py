"""ctypes binding of libhhsr.so — the only way this package reaches the GPU for its own arithmetic.

There is deliberately no fallback: if the shared library is missing (not built) the import fails loudly; a
non-zero status from any entry point raises RuntimeError with the library's message.  Signatures follow
include/hhsr.h exactly (plain pointers and sizes, no torch types cross the boundary).
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HHSR_LIB", os.path.join(_HERE, "libhhsr.so"))   # HHSR_LIB: kernel-variant experiments

_P, _I, _D, _F, _Z = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_size_t
_IP, _DP, _FP = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)

# name -> argtypes (every function returns int except the two noted); mirrors include/hhsr.h
SIGNATURES = {
    "hhsr_grey_band_mask": [_P, _I, _I, C.c_longlong, C.c_longlong, _F, _P],
    "hhsr_grey_fft_sizes": [_I, _I, C.POINTER(_Z), C.POINTER(_Z)],
    "hhsr_grey_fft_plan": [_P, _I, _I, _P],
    "hhsr_grey_fft": [_P, _I, _I, _P, _P, _P, _P],
    "hhsr_pad_circular": [_P, _I, _I, _P, _I, _I, _P],
    "hhsr_gauss_downsample": [_P, _I, _I, _I, _FP, _I, _P, _I, _I, _P],
    "hhsr_grad_hessian": [_P, _I, _I, _I, _P, _P, _P, _P],
    "hhsr_upscale_flow": [_P, _I, _I, _P, _I, _I, _I, _F, _I, _P],
    "hhsr_bm_l2_search": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _P],
    "hhsr_bm_l1_compat": [_P, _I, _P],
    "hhsr_bm_l1_search": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _P],
    "hhsr_ica": [_P, _P, _P, _I, _I, _P, _P, _I, _I, _P, _I, _I, _I, _I, _P],
    "hhsr_estimate_kernels": [_P, _I, _I, _D, _D, _D, _D, _D, _D, _D, _D, _I, _P, _P],
    "hhsr_gat": [_P, _Z, _D, _D, _P, _P],
    "hhsr_decimate_to_grey": [_P, _I, _I, _P, _P],
    "hhsr_guide_stats": [_P, _I, _I, _IP, _DP, _P, _P, _P],
    "hhsr_guide_image": [_P, _I, _I, _IP, _DP, _P, _P],
    "hhsr_local_stats": [_P, _I, _I, _I, _P, _P, _P],
    "hhsr_upscale_warp_stats": [_P, _I, _I, _P, _I, _I, _I, _P, _P],
    "hhsr_noise_table": [_P, _P, _I, _P, _P],
    "hhsr_ref_stats_terms": [_P, _P, _I, _I, _P, _I, _P, _P, _P, _P],
    "hhsr_robustness_ref_terms": [_P, _P, _I, _I, _P, _I, _P, _P],
    "hhsr_robustness": [_P, _P, _P, _I, _I, _P, _I, _I, _I, _D, _D, _D, _D, _P, _I, _P],
    "hhsr_local_min5": [_P, _I, _I, _P, _P, _P],
    "hhsr_merge_accumulate": [_P, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _D, _IP, _I, _P],
    "hhsr_merge_init_accumulate": [_P, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _D, _IP, _I, _P],
    "hhsr_merge_accumulate_batch": [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _I, _I,
                                    _I, _P, _P, _I, _I, _D, _IP, _I, _I, _P],
    "hhsr_merge_ref": [_P, _I, _I, _P, _P, _P, _I, _I, _D, _IP, _I, _P, _I, _I, _D, _I, _I, _I, _P],
    "hhsr_merge_accumulate_rows": [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _I, _I,
                                   _I, _P, _P, _I, _I, _D, _IP, _I, _I, _I, _I, _P],
    "hhsr_merge_finish_rows": [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _I, _I,
                               _I, _P, _P, _I, _I, _D, _IP, _I, _I, _I, _I, _P, _P, _P, _P],
    "hhsr_merge_ref_rows": [_P, _I, _I, _P, _P, _P, _I, _I, _D, _IP, _I, _P, _I, _I, _D, _I, _I, _I, _P],
    "hhsr_gather_bands": [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                          C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "hhsr_normalize_raw_u16": [_P, _I, _I, _FP, _FP, _FP, _P, _P],
    "hhsr_reduce_merge_ref": [C.POINTER(_P), C.POINTER(_P), _I, _P, _I, _I, _P, _P, _I, _I, _D, _IP, _I, _P, _I, _I, _D, _I,
                              _I, _I, _P],
    "hhsr_post_ccm_clip": [_P, _Z, _FP, _P],
    "hhsr_post_blur_cols": [_P, _I, _I, _DP, _I, _P, _P],
    "hhsr_post_finish": [_P, _P, _I, _I, _DP, _I, _F, _I, _F, _I, _P, _P],
    "hhsr_frame_count_denoise_gauss": [_P, _I, _I, _P, _I, _I, _D, _D, _D, _P, _P],
    "hhsr_frame_count_denoise_median": [_P, _I, _I, _P, _I, _I, _D, _D, _D, _P, _P],
    "hhsr_noise_mc": [_P, _I, _D, _D, _I, C.c_ulonglong, _P, _P, _P],
    "hhsr_divide": [_P, _P, _Z, _P],
    "hhsr_add_f64_f32": [_P, _P, _Z, _P],
    "hhsr_add_many_f64_f32": [_P, C.POINTER(_P), _I, _Z, _P],
}

_lib = None
launch_count = 0   # number of libhhsr kernels enqueued by this process (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libhhsr.so is not built (%s missing): run `python __graft_entry__.py build` or `make -C "
                "handheld-multi-frame-super-resolution_b200/csrc`.  There is no CPU or PyTorch fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hhsr_version.restype = C.c_int
        L.hhsr_last_error_string.restype = C.c_char_p
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def call(name, *args):
    """Invoke an entry point; raise RuntimeError on a non-zero status (include/hhsr.h error convention)."""
    global launch_count
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, L.hhsr_last_error_string().decode()))
    launch_count += 1


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def as_device(x, dtype=torch.float32):
    """Anything array-like (numpy, torch, __cuda_array_interface__) -> contiguous CUDA tensor of `dtype`.
    The reference mixes Numba device arrays and torch tensors freely at this boundary (e.g. ICA.py:23-24)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


_HOST_ARRAYS = {}     # small host-side argument arrays, cached by value (they are rebuilt for every launch otherwise)


def _cached(kind, values, build):
    key = (kind, values)
    arr = _HOST_ARRAYS.get(key)
    if arr is None:
        if len(_HOST_ARRAYS) > 256:
            _HOST_ARRAYS.clear()
        arr = _HOST_ARRAYS[key] = build(values)
    return arr


def cfa_array(cfa):
    a = np.asarray(cfa.cpu() if isinstance(cfa, torch.Tensor) else cfa).astype(np.int64).reshape(-1)
    if a.size != 4:
        raise ValueError("CFA pattern must be 2x2")
    return _cached("cfa", tuple(int(v) for v in a), lambda v: (C.c_int * 4)(*v))


def wb_array(wb):
    a = np.asarray(wb.cpu() if isinstance(wb, torch.Tensor) else wb, dtype=np.float64).reshape(-1)
    if a.size < 3:
        raise ValueError("white balance needs at least 3 gains")
    return _cached("wb", tuple(float(v) for v in a[:3]), lambda v: (C.c_double * 3)(*v))
