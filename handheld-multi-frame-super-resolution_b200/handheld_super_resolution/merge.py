"""Kernel-regression merge (Alg. 4, Alg. 11) — mirrors handheld_super_resolution/merge.py of the reference
(merge :236-288, merge_ref :22-80).  num / den [Hs, Ws, 3] float32 are accumulated IN PLACE."""
import ctypes as C

import torch

from . import _lib


def _check_acc(num, den):
    assert num.shape == den.shape and num.shape[-1] == 3
    assert num.is_cuda and den.is_cuda and num.is_contiguous() and den.is_contiguous()
    assert num.dtype == torch.float32 and den.dtype == torch.float32


def merge(comp_img, alignments, covs, r, num, den, cfa_pattern, config, init=False):
    """Accumulate comp frame J_n into num/den with its flow, covariances and robustness (merge.py:236-288).
    init=True (B200 addition): num/den are initialised with this frame's contribution — equal to accumulating into
    zero-filled arrays, without needing them zero-filled (main() does this for the first comp frame)."""
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    _check_acc(num, den)
    comp_img = _lib.as_device(comp_img)
    H, W = comp_img.shape
    iso = config.merging.kernel == "iso"
    ny, nx, _ = alignments.shape
    _lib.call("hhsr_merge_init_accumulate" if init else "hhsr_merge_accumulate", _lib.ptr(comp_img), H, W, _lib.ptr(alignments), ny, nx,
              int(config.block_matching.tuning.tile_size), _lib.ptr(None if iso else covs), _lib.ptr(r), _lib.ptr(num),
              _lib.ptr(den), num.shape[0], num.shape[1], float(config.scale), _lib.cfa_array(cfa_pattern), int(iso),
              _lib.stream())


MERGE_INIT, MERGE_GENERIC = 1, 2      # include/hhsr.h: HHSR_MERGE_INIT, HHSR_MERGE_GENERIC


def fast_path_applies(H, W, scale, tile_size):
    """Whether the power-of-two merge kernels run (scales 1, 2, 4; output width a multiple of 4; power-of-two tile size
    >= 4) — the condition of csrc/merge.cu pow2_fast_shift(), needed on the host to plan the fused finish."""
    ts = int(tile_size)
    return scale in (1, 2, 4) and (W * int(scale)) % 4 == 0 and ts >= 4 and ts & (ts - 1) == 0


def merge_batch(comp_imgs, alignments, covs, rs, num, den, cfa_pattern, config, init=False, generic=False, finish=None):
    """Same arithmetic as calling merge() once per frame in list order (bit-identical), in ONE pass over num/den: the
    accumulators are read and written once instead of once per frame (B200 addition, SURVEY section 8d "K-frame
    batched merge").  init=True: the batch initialises num/den (their previous contents are ignored) — what main() does
    with the first batch of a burst.  generic=True forces the any-scale kernel (A/B parity tests).
    finish=(ref_img, ref_kernels): this is the LAST batch of the burst — merge_ref (plain mode) and utils.divide follow in
    the same pass on the register accumulators and `num` receives the finished image (bit-identical to merge_batch +
    merge_ref(fuse_divide=True)); `den` is left untouched.  Needs fast_path_applies()."""
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    _check_acc(num, den)
    K = len(comp_imgs)
    H, W = comp_imgs[0].shape
    iso = config.merging.kernel == "iso"
    ny, nx, _ = alignments[0].shape
    arr = lambda ts: (C.c_void_p * K)(*[t.data_ptr() if t is not None else 0 for t in ts])  # noqa: E731
    if finish is not None:
        ref_img, ref_kernels = finish
        _lib.call("hhsr_merge_finish_rows", arr(comp_imgs), arr(alignments), arr([None] * K if iso else covs), arr(rs),
                  K, H, W, ny, nx, int(config.block_matching.tuning.tile_size), _lib.ptr(num), _lib.ptr(den), num.shape[0],
                  num.shape[1], float(config.scale), _lib.cfa_array(cfa_pattern), int(iso), MERGE_INIT if init else 0,
                  0, num.shape[0], _lib.ptr(ref_img), _lib.ptr(None if iso else ref_kernels), _lib.ptr(num), _lib.stream())
        return
    _lib.call("hhsr_merge_accumulate_batch", arr(comp_imgs), arr(alignments), arr([None] * K if iso else covs), arr(rs),
              K, H, W, ny, nx, int(config.block_matching.tuning.tile_size), _lib.ptr(num), _lib.ptr(den), num.shape[0],
              num.shape[1], float(config.scale), _lib.cfa_array(cfa_pattern), int(iso),
              (MERGE_INIT if init else 0) | (MERGE_GENERIC if generic else 0), _lib.stream())


def merge_ref(ref_img, kernels, num, den, cfa_pattern, config, acc_rob=None, fuse_divide=False, rows=None):
    """Accumulate the reference frame (merge.py:22-80).  With the accumulated-robustness denoiser enabled,
    `acc_rob` (float64 [H,W]) widens the window / overwrites the accumulators where few frames were merged.
    fuse_divide=True also performs utils.divide(num, den) in the same pass; rows=(begin, end) restricts the work
    to a slice of output rows (both B200 additions, used by the frame-sharded driver)."""
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    _check_acc(num, den)
    ref_img = _lib.as_device(ref_img)
    H, W = ref_img.shape
    iso = config.merging.kernel == "iso"
    ard = config.accumulated_robustness_denoiser
    if ard.enabled:
        if acc_rob is None:
            raise ValueError("accumulated robustness denoiser enabled but no acc_rob given")
        acc = _lib.as_device(acc_rob, torch.float64)
        rad_max, max_mult, max_fc = int(ard.merge.rad_max), float(ard.merge.max_multiplier), int(ard.merge.max_frame_count)
    else:
        acc, rad_max, max_mult, max_fc = None, 0, 0.0, 0
    _lib.call("hhsr_merge_ref", _lib.ptr(ref_img), H, W, _lib.ptr(None if iso else kernels), _lib.ptr(num), _lib.ptr(den),
              num.shape[0], num.shape[1], float(config.scale), _lib.cfa_array(cfa_pattern), int(iso), _lib.ptr(acc),
              max_fc, rad_max, max_mult, int(fuse_divide), *(rows if rows is not None else (0, num.shape[0])), _lib.stream())
