"""Block matching — mirrors handheld_super_resolution/block_matching.py of the reference
(align_lvl_block_matching_L2 :20-76, align_lvl_block_matching_L1 :78-103)."""
import torch

from . import _lib


def align_lvl_block_matching_L2(tyled_pyr_lvl, ref_fft_lvl, moving_lvl, alignment, l, config):
    """Exhaustive L2 search around rint(alignment), updating `alignment` IN PLACE (block_matching.py:20-76).

    `tyled_pyr_lvl` is the reference pyramid level itself and `ref_fft_lvl` is unused: the reference correlates
    through batched FFTs of pre-padded tiles, this implementation searches the SSD directly in shared memory
    (same function), so init_alignment hands over the plain level instead of tiles + spectra."""
    ts = config.block_matching.tuning.tile_sizes[l]
    radius = config.block_matching.tuning.search_radii[l]
    if ts not in (8, 16, 32, 64):
        raise NotImplementedError("Box filter for tile size {} not implemented".format(ts))
    assert alignment.is_cuda and alignment.is_contiguous() and alignment.dtype == torch.float32
    ny, nx, _ = alignment.shape
    rh, rw = tyled_pyr_lvl.shape
    mh, mw = moving_lvl.shape
    _lib.call("hhsr_bm_l2_search", _lib.ptr(tyled_pyr_lvl), rh, rw, _lib.ptr(moving_lvl), mh, mw, _lib.ptr(alignment),
              ny, nx, int(ts), int(radius), _lib.stream())


_WARNED = set()


def align_lvl_block_matching_L1(ref_lvl, moving_lvl, alignments, l, config):
    """The L1 level (block_matching.py:78-345).

    Default (`block_matching.tuning.l1_compat` absent or true): what the compiled reference EXECUTES — its SAD search
    never updates the shift (`if err < min` with err = +inf), so the net effect is alignments <- rint(alignments)
    (confirmed on B200 for tile sizes 32 and 64 by baseline/probe_reference.py).  Tile size 16 is undefined
    behaviour upstream (a data race) and is given the same well-defined semantics here.  A one-time warning
    says that search_radii[l] is ignored on this path.

    `l1_compat: false`: what the reference INTENDS — an exhaustive sum |ref - moving| search over (2r+1)^2 shifts around
    rint(alignment), zero outside the frame, first minimum (hhsr_bm_l1_search).  Not parity with upstream's outputs."""
    tuning = config.block_matching.tuning
    ts = tuning.tile_sizes[l]
    radius = tuning.search_radii[l]
    if ts not in (16, 32, 64):
        raise NotImplementedError("L1 local search kernel for tile size {} not implemented".format(ts))
    if ts == 16:
        assert 2 * radius + 16 <= 32, "L1 local search kernel only implemented for search windows up to size 32"
    if ts == 64:
        assert 2 * radius <= 16, f"Cant handle search radius {radius} with tile size {ts} in L1 local search kernel."
    assert alignments.is_cuda and alignments.is_contiguous() and alignments.dtype == torch.float32
    compat = tuning.get("l1_compat", True) if hasattr(tuning, "get") else getattr(tuning, "l1_compat", True)
    if compat:
        if radius > 0 and "l1" not in _WARNED:
            _WARNED.add("l1")
            import warnings
            warnings.warn("L1 block-matching level: reproducing the compiled reference, which does not search (flow <- "
                          "rint(flow), search radius ignored); set block_matching.tuning.l1_compat = false for the intended "
                          "SAD search", stacklevel=2)
        _lib.call("hhsr_bm_l1_compat", _lib.ptr(alignments), alignments.numel(), _lib.stream())
        return
    ny, nx, _ = alignments.shape
    rh, rw = ref_lvl.shape
    mh, mw = moving_lvl.shape
    _lib.call("hhsr_bm_l1_search", _lib.ptr(ref_lvl), rh, rw, _lib.ptr(moving_lvl), mh, mw, _lib.ptr(alignments),
              ny, nx, int(ts), int(radius), _lib.stream())
