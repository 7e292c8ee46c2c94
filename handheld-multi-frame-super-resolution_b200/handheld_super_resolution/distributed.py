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


def main_sharded(ref_img, comp_imgs, config, group=None):
    """main() with the comp frames of this rank only and one all-reduce at the reduction point.  Every rank
    returns the full normalised image (identical up to float32 summation order)."""
    from .super_resolution import main
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    ids = shard_frames(len(comp_imgs), rank, world)
    return main(ref_img, comp_imgs, config, frame_ids=ids,
                reduce_fn=lambda n, d, a: allreduce_accumulators(n, d, a, group))
