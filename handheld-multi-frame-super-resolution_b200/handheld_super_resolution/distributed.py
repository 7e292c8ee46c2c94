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


def main_sharded(ref_img, comp_imgs, config, group=None, mode=None):
    """main() with the comp frames of this rank only and one sum of the accumulators at the reduction point
    (mode "reduce_scatter", default, or "allreduce"; env HHSR_SHARD_REDUCE overrides).  Every rank returns the full
    normalised image (identical up to float32 summation order)."""
    import os
    from .super_resolution import main
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mode = mode or os.environ.get("HHSR_SHARD_REDUCE", "reduce_scatter")
    ids = shard_frames(len(comp_imgs), rank, world)
    if mode == "allreduce":
        fn = lambda n, d, a: allreduce_accumulators(n, d, a, group)   # noqa: E731
    else:
        fn = lambda n, d, a: reduce_scatter_accumulators(n, d, a, group)   # noqa: E731
    return main(ref_img, comp_imgs, config, frame_ids=ids, reduce_fn=fn)
