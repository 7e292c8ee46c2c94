"""Frame-sharded multi-GPU driver — a B200 addition, the reference is single-GPU.

One process per GPU (torchrun).  Comp frames are dealt round-robin to ranks; every rank recomputes the
reference-side products (one frame's worth of work, no communication) and accumulates its frames into private
num/den/acc_rob.  The ONLY collective of the pipeline is one sum of those accumulators over NCCL (NVLink 5 /
NVSwitch) after the frame loop; merge_ref + divide then run on the reduced accumulators."""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world_size):
    """Frames {i : i mod G == rank}: 19 frames over 8 ranks -> 3,3,3,2,2,2,2,2."""
    return list(range(rank, n_frames, world_size))


def allreduce_accumulators(num, den, acc_rob=None, group=None):
    """The one reduction point: element-wise float32 sum of num and den (float64 for acc_rob) across ranks.
    num and den are views of one flat buffer when allocated by main_sharded, so this is a single large message."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    if num.untyped_storage().data_ptr() == den.untyped_storage().data_ptr():
        flat = torch.empty(0, dtype=num.dtype, device=num.device).set_(num.untyped_storage())
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)
    if acc_rob is not None:
        dist.all_reduce(acc_rob, op=dist.ReduceOp.SUM, group=group)


def reduce_scatter_accumulators(num, den, acc_rob=None, group=None):
    """Same reduction point, cheaper data movement: the sum is delivered as a REDUCE-SCATTER by slices of output
    rows (each rank receives the summed num/den of its own slice, in place), the caller normalises only that slice
    (merge_ref + divide on 1/G of the image) and `gather(num)` re-assembles the finished image on every rank with
    an all-gather.  Falls back to allreduce_accumulators when the row count does not divide by the world size.
    Returns ((row_begin, row_end), gather) or None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Hs = num.shape[0]
    if Hs % world != 0 or not (num.is_contiguous() and den.is_contiguous()):
        allreduce_accumulators(num, den, acc_rob, group)
        return None
    rows = Hs // world
    nf, df = num.view(-1), den.view(-1)
    chunk = nf.numel() // world
    dist.reduce_scatter_tensor(nf[rank * chunk:(rank + 1) * chunk], nf, op=dist.ReduceOp.SUM, group=group)
    dist.reduce_scatter_tensor(df[rank * chunk:(rank + 1) * chunk], df, op=dist.ReduceOp.SUM, group=group)
    if acc_rob is not None:
        dist.all_reduce(acc_rob, op=dist.ReduceOp.SUM, group=group)

    def gather(image):
        flat = image.view(-1)
        dist.all_gather_into_tensor(flat, flat[rank * chunk:(rank + 1) * chunk], group=group)
    return (rank * rows, (rank + 1) * rows), gather


def p2p_row_slices(Hs, world):
    """Row slices [(begin, end)] of the fused peer-memory reduction.  Rank 0 also RECEIVES every other rank's finished
    slice (half the bytes of the num + den it would otherwise pull), so it takes a slice half as tall: inbound NVLink
    bytes are then equal on all ranks ((G-1)(a S + b S/2) = (G-1) b S  =>  a = b/2)."""
    w = [0.5] + [1.0] * (world - 1)
    tot = sum(w)
    edge = [0.0]
    for x in w:
        edge.append(edge[-1] + x)
    cuts = [round(e / tot * Hs) for e in edge]
    cuts[-1] = Hs
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class P2PReduce:
    """The reduction point as ONE kernel over NVLink peer memory (mode "p2p"): every rank keeps its private num/den in
    a symmetric-memory buffer (torch.distributed._symmetric_memory: the same allocation mapped into every rank's
    address space); after a device-side barrier rank g runs hhsr_reduce_merge_ref on its slice of output rows — it
    pulls the G partial accumulators of that slice straight from the peers' HBM, adds the reference frame, divides,
    and stores the finished pixels into rank 0's buffer.  No NCCL pass, no separate merge_ref / divide pass, and the
    1/G slices travel once.  Only rank 0 ends up with the whole image (a gather, not an all-gather).
    Buffers and the rendezvous (~2 s) are cached per output shape."""
    _cache = {}

    def __init__(self, shape, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.shape = tuple(shape)
        n = 1
        for d in self.shape:
            n *= d
        self.numel = n
        self.flat = symm_mem.empty((2 * n,), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm_mem.rendezvous(self.flat, self.group)
        self.num = self.flat[:n].view(self.shape)
        self.den = self.flat[n:].view(self.shape)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    @classmethod
    def get(cls, shape, group=None):
        key = (tuple(shape), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(shape, group)
        return cls._cache[key]

    def rows(self):
        return p2p_row_slices(self.shape[0], self.world)[self.rank]

    def finalize(self, ref_img, covs, num, den, acc_rob, cfa_pattern, config):
        import ctypes as C
        from . import _lib
        assert num.data_ptr() == self.num.data_ptr() and den.data_ptr() == self.den.data_ptr()
        ard = config.accumulated_robustness_denoiser
        if acc_rob is not None:
            dist.all_reduce(acc_rob, op=dist.ReduceOp.SUM, group=self.group)
        if ard.enabled:
            acc, rad_max, max_mult, max_fc = acc_rob, int(ard.merge.rad_max), float(ard.merge.max_multiplier), int(ard.merge.max_frame_count)
        else:
            acc, rad_max, max_mult, max_fc = None, 0, 0.0, 0
        self.hdl.barrier(channel=0)                 # every rank has finished accumulating its frames
        H, W = ref_img.shape
        iso = config.merging.kernel == "iso"
        nums = (C.c_void_p * self.world)(*self.peer_ptrs)
        dens = (C.c_void_p * self.world)(*[p + 4 * self.numel for p in self.peer_ptrs])
        r0, r1 = self.rows()
        _lib.call("hhsr_reduce_merge_ref", nums, dens, self.world, _lib.ptr(ref_img), H, W, _lib.ptr(None if iso else covs),
                  C.c_void_p(self.peer_ptrs[0]), self.shape[0], self.shape[1], float(config.scale),
                  _lib.cfa_array(cfa_pattern), int(iso), _lib.ptr(acc), max_fc, rad_max, max_mult, 1, r0, r1, _lib.stream())
        self.hdl.barrier(channel=1)                 # all slices delivered; peers may reuse their accumulators
        return num                                  # the whole image on rank 0 only


def main_sharded(ref_img, comp_imgs, config, group=None, mode=None):
    """main() with the comp frames of this rank only and one sum of the accumulators at the reduction point
    (mode "reduce_scatter", default, "allreduce", or "p2p" — the fused peer-memory kernel, see P2PReduce; env
    HHSR_SHARD_REDUCE overrides).  With the NCCL modes every rank returns the full normalised image (identical up to
    float32 summation order); with "p2p" only rank 0 does."""
    import os
    from .super_resolution import main
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mode = mode or os.environ.get("HHSR_SHARD_REDUCE", "reduce_scatter")
    ids = shard_frames(len(comp_imgs), rank, world)
    if mode == "p2p" and world > 1:
        H, W = ref_img.shape
        s = config.scale
        red = P2PReduce.get((round(s * H), round(s * W), 3), group)
        return main(ref_img, comp_imgs, config, frame_ids=ids, accumulators=(red.num, red.den), finalize_fn=red.finalize)
    if mode == "allreduce":
        fn = lambda n, d, a: allreduce_accumulators(n, d, a, group)   # noqa: E731
    else:
        fn = lambda n, d, a: reduce_scatter_accumulators(n, d, a, group)   # noqa: E731
    return main(ref_img, comp_imgs, config, frame_ids=ids, reduce_fn=fn)
