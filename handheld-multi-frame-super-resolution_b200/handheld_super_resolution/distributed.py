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
        try:
            if r1 > r0:       # an empty slice (tiny images) has nothing to launch but must still reach the second barrier
                _lib.call("hhsr_reduce_merge_ref", nums, dens, self.world, _lib.ptr(ref_img), H, W, _lib.ptr(None if iso else covs),
                          C.c_void_p(self.peer_ptrs[0]), self.shape[0], self.shape[1], float(config.scale),
                          _lib.cfa_array(cfa_pattern), int(iso), _lib.ptr(acc), max_fc, rad_max, max_mult, 1, r0, r1, _lib.stream())
        finally:
            self.hdl.barrier(channel=1)                 # all slices delivered; peers may reuse their accumulators
        # the symmetric accumulators are overwritten by the next burst: hand out a private copy (rank 0 holds the whole image)
        return num.clone() if self.rank == 0 else num


def equal_row_slices(Hs, world, multiple=8):
    """Output-row slices [(begin, end)] of the row-sharded merge: as equal as possible, cut at multiples of `multiple`
    rows (the merge kernels work on blocks of 8 rows), never empty while Hs >= world * multiple."""
    blocks = -(-Hs // multiple)
    cuts = [min(Hs, multiple * ((blocks * r) // world)) for r in range(world)] + [Hs]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class RowShardedMerge:
    """Mode "rows": frames are sharded over ranks for alignment, robustness and kernel estimation; the MERGE is sharded by
    OUTPUT ROWS.  Each rank publishes the products of its frames (raw frame, flow, covariances, robustness: 144 MB per
    12 MP frame) in symmetric memory (torch.distributed._symmetric_memory: one allocation mapped into every rank over
    NVLink / NVSwitch).  After ONE device-side barrier — the only exchange point of the pipeline — a rank pulls, for
    every frame of the burst, just the LR row band its slice of output rows can touch (hhsr_gather_bands, extents derived
    on the device from the flow) and merges ALL frames in one pass with the accumulators in registers
    (hhsr_merge_accumulate_rows), then adds the reference frame and divides (hhsr_merge_ref_rows).

    Against the frame-sharded sum of accumulators ("p2p" / "reduce_scatter"): per rank ~0.3 GB instead of 1.0 GB cross
    NVLink at 8 GPUs for a 20 x 12 MP burst, nothing is summed across ranks, the merge work is balanced whatever the
    number of frames per rank, and — frames being accumulated in burst order by the same kernel — the image is
    BIT-IDENTICAL to the single-GPU result.  Each rank ends with its own slice of the image ([rows, Ws, 3]); nothing is
    gathered (the host-to-host path copies the slices over the ranks' own PCIe links)."""
    _cache = {}

    def __init__(self, H, W, scale, n_comp, ts, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.H, self.W, self.scale, self.n_comp, self.ts = H, W, scale, n_comp, ts
        self.Hs, self.Ws = round(scale * H), round(scale * W)
        if self.Ws % 4 != 0 or W % 4 != 0 or H % 2 != 0:
            raise ValueError("row-sharded merge needs W and the output width to be multiples of 4")
        self.ny, self.nx = -(-H // ts), -(-W // ts)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.slots = max(1, -(-n_comp // self.world))
        # per slot: raw [H,W] | r [H,W] | covs [H/2,W/2,4] | flow [ny,nx,2] (padded to 4 floats)
        self.n_plane, self.n_flow = H * W, -(-(self.ny * self.nx * 2) // 4) * 4
        self.slot_floats = 3 * self.n_plane + self.n_flow
        self.sym = symm_mem.empty((self.slots * self.slot_floats,), dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.sym, self.group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # local full-size planes of every frame of the burst (only the gathered bands are ever written / read)
        self.local = torch.empty((n_comp * self.slot_floats,), dtype=torch.float32, device=dev)
        self.extents = torch.zeros((max(n_comp, 1), 4), dtype=torch.int32, device=dev)
        self.row_slice = equal_row_slices(self.Hs, self.world)[self.rank]
        rows = self.row_slice[1] - self.row_slice[0]
        self.num = torch.empty((rows, self.Ws, 3), dtype=torch.float32, device=dev)
        self.den = torch.empty((rows, self.Ws, 3), dtype=torch.float32, device=dev)
        self.n_local = 0

    @classmethod
    def get(cls, H, W, scale, n_comp, ts, group=None):
        key = (H, W, scale, n_comp, ts, id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(H, W, scale, n_comp, ts, group)
        return cls._cache[key]

    def _slot_views(self, flat, j):
        b = flat[j * self.slot_floats:(j + 1) * self.slot_floats]
        n, H, W = self.n_plane, self.H, self.W
        return (b[:n].view(H, W), b[n:2 * n].view(H, W), b[2 * n:3 * n].view(H // 2, W // 2, 2, 2),
                b[3 * n:3 * n + self.ny * self.nx * 2].view(self.ny, self.nx, 2))

    def outputs(self, k):
        """Where main() writes the robustness and covariances of this rank's k-th frame: straight into its slot."""
        _, r_s, covs_s, _ = self._slot_views(self.sym, k)
        return r_s, covs_s

    def publish(self, k, im_id, frame, flow, covs, r):
        """Publish the products of this rank's k-th frame (burst index im_id, owner im_id % world, slot im_id // world)."""
        assert im_id % self.world == self.rank and im_id // self.world == k < self.slots
        raw_s, r_s, covs_s, flow_s = self._slot_views(self.sym, k)
        raw_s.copy_(frame), flow_s.copy_(flow)
        if r.data_ptr() != r_s.data_ptr():
            r_s.copy_(r)
        if covs is not None and covs.data_ptr() != covs_s.data_ptr():
            covs_s.copy_(covs)
        self.n_local = k + 1

    def finalize(self, ref_img, covs_ref, num, den, acc_rob, cfa_pattern, config):
        import ctypes as C
        from . import _lib
        from .utils import add_many
        n, world = self.n_comp, self.world
        iso = config.merging.kernel == "iso"
        r0, r1 = self.row_slice
        rows = r1 - r0
        out = torch.empty((rows, self.Ws, 3), dtype=torch.float32, device=self.num.device) if rows > 0 else None
        from .super_resolution import _mark
        self.hdl.barrier(channel=0)                 # every rank has published its frames: THE exchange point
        _mark("rows_barrier")
        try:
            if rows > 0 and n > 0:
                F = self.slot_floats * 4
                src = [self.peer_ptrs[f % world] + (f // world) * F for f in range(n)]
                own = [f % world == self.rank for f in range(n)]
                dst = [src[f] if own[f] else self.local.data_ptr() + f * F for f in range(n)]
                P = 4 * self.n_plane
                arr = lambda xs: (C.c_void_p * n)(*xs)  # noqa: E731
                lr0 = int(r0 / self.scale)
                lr1 = min(self.H, int(-(-r1 // self.scale)) + 1)
                _lib.call("hhsr_gather_bands", arr(src), arr([p + P for p in src]), arr([0 if iso else p + 2 * P for p in src]),
                          arr([p + 3 * P for p in src]), arr(dst), arr([p + P for p in dst]),
                          arr([0 if iso else p + 2 * P for p in dst]), arr([p + 3 * P for p in dst]), n, self.H, self.W,
                          self.ny, self.nx, int(self.ts), lr0, lr1, _lib.ptr(self.extents), _lib.stream())
                _mark("rows_gather")
                from .merge import fast_path_applies
                fused = not config.accumulated_robustness_denoiser.enabled and fast_path_applies(self.H, self.W, self.scale, self.ts)
                margs = (arr(dst), arr([p + 3 * P for p in dst]), arr([0 if iso else p + 2 * P for p in dst]),
                         arr([p + P for p in dst]), n, self.H, self.W, self.ny, self.nx, int(self.ts), _lib.ptr(self.num),
                         _lib.ptr(self.den), self.Hs, self.Ws, float(self.scale), _lib.cfa_array(cfa_pattern), int(iso), 1, r0, r1)
                if fused:      # all frames + the reference frame + divide in ONE pass: only the finished slice is written
                    _lib.call("hhsr_merge_finish_rows", *margs, _lib.ptr(ref_img), _lib.ptr(None if iso else covs_ref),
                              _lib.ptr(out), _lib.stream())
                else:
                    _lib.call("hhsr_merge_accumulate_rows", *margs, _lib.stream())
                _mark("rows_merge")
                if acc_rob is not None:
                    # accumulated robustness of the LR rows [lr0, lr1) this slice looks at (merge_ref reads it at
                    # rint(oy / scale)): the sum over ALL frames in burst order, from the gathered bands; the other rows
                    # keep this rank's frames only
                    def r_rows(f):
                        flat, j = (self.sym, f // world) if own[f] else (self.local, f)
                        start = j * self.slot_floats + self.n_plane + lr0 * self.W
                        return flat[start:start + (lr1 - lr0) * self.W].view(lr1 - lr0, self.W)
                    acc_rob[lr0:lr1].zero_()
                    add_many(acc_rob[lr0:lr1], [r_rows(f) for f in range(n)])
            elif rows > 0:
                self.num.zero_(), self.den.zero_()
                fused = False
            if rows > 0 and not fused:
                ard = config.accumulated_robustness_denoiser
                if ard.enabled:
                    acc, rad_max, max_mult, max_fc = acc_rob, int(ard.merge.rad_max), float(ard.merge.max_multiplier), int(ard.merge.max_frame_count)
                else:
                    acc, rad_max, max_mult, max_fc = None, 0, 0.0, 0
                _lib.call("hhsr_merge_ref_rows", _lib.ptr(ref_img), self.H, self.W, _lib.ptr(None if iso else covs_ref),
                          _lib.ptr(self.num), _lib.ptr(self.den), self.Hs, self.Ws, float(self.scale), _lib.cfa_array(cfa_pattern),
                          int(iso), _lib.ptr(acc), max_fc, rad_max, max_mult, 1, r0, r1, _lib.stream())
                out.copy_(self.num)
            _mark("rows_merge_ref")
        finally:
            self.hdl.barrier(channel=1)             # all bands pulled: the owners may overwrite their slots (next burst)
        return out


def broadcast_reference_frame(ref_img, config, group=None):
    """Every rank needs the reference frame.  When it lives in host memory, rank 0 uploads it once and NCCL broadcasts it
    over NVLink (48 MB at 12 MP) instead of every rank pulling its own copy through the host's PCIe complex — with 8
    ranks that is 7 frames less of host-to-device traffic per burst.  Input distribution only: the merge path still has
    its single exchange point.  Device-resident frames pass through.

    Rank 0 stages the frame like main() stages every host frame — on the copy stream, through the persistent staging ring
    (FrameFeeder), so the upload of burst i + 1's reference frame overlaps the compute of burst i — and broadcasts the
    NORMALISED float32 frame (uint16 sensor counts are normalised on rank 0 only; the other ranks receive exactly what
    rank 0 computes with)."""
    if isinstance(ref_img, torch.Tensor) and ref_img.is_cuda:
        return ref_img
    from .super_resolution import FrameFeeder, _host_tensor
    dev = torch.device("cuda", torch.cuda.current_device())
    host = _host_tensor(ref_img)
    if dist.get_rank(group) == 0:
        feed = FrameFeeder([host], [0], config, dev, role="ref")
        buf = feed.get(0).clone()          # the frame outlives its staging slot
        feed.release(0)
    else:
        buf = torch.empty(host.shape, dtype=torch.float32, device=dev)
    dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return buf


def main_sharded(ref_img, comp_imgs, config, group=None, mode=None):
    """main() with the comp frames of this rank only and ONE exchange point after the frame loop:
      "reduce_scatter" (default) / "allreduce": the frame-sharded accumulators are summed over NCCL; every rank returns
          the full normalised image (identical up to float32 summation order);
      "p2p": the same sum as one fused peer-memory kernel (P2PReduce); only rank 0 returns the whole image;
      "rows": the merge itself is sharded by output rows (RowShardedMerge): every rank returns ITS slice
          [rows, Ws, 3] of the image, debug_dict["rows"] = (begin, end); bit-identical to the single-GPU image.
    env HHSR_SHARD_REDUCE overrides the default."""
    import os
    from .super_resolution import main
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if world == 1:
        return main(ref_img, comp_imgs, config)       # nothing to exchange: the plain single-GPU path (fused finish included)
    mode = mode or os.environ.get("HHSR_SHARD_REDUCE", "reduce_scatter")
    ids = shard_frames(len(comp_imgs), rank, world)
    ahead = None
    if world > 1:
        ref_img = broadcast_reference_frame(ref_img, config, group)
        ahead = max(1, min(len(ids), int(os.environ.get("HHSR_SHARD_ALIGN_AHEAD", "3"))))   # few frames per rank: run their chains together
    if mode == "rows" and world > 1:
        H, W = ref_img.shape
        rs = RowShardedMerge.get(H, W, config.scale, len(comp_imgs), int(config.block_matching.tuning.tile_size), group)
        out, dbg = main(ref_img, comp_imgs, config, frame_ids=ids, frame_sink=rs, finalize_fn=rs.finalize, align_ahead=ahead)
        dbg["rows"] = rs.row_slice
        return out, dbg
    if mode == "p2p" and world > 1:
        H, W = ref_img.shape
        s = config.scale
        red = P2PReduce.get((round(s * H), round(s * W), 3), group)
        return main(ref_img, comp_imgs, config, frame_ids=ids, accumulators=(red.num, red.den), finalize_fn=red.finalize,
                    align_ahead=ahead)
    if mode == "allreduce":
        fn = lambda n, d, a: allreduce_accumulators(n, d, a, group)   # noqa: E731
    else:
        fn = lambda n, d, a: reduce_scatter_accumulators(n, d, a, group)   # noqa: E731
    return main(ref_img, comp_imgs, config, frame_ids=ids, reduce_fn=fn, align_ahead=ahead)
