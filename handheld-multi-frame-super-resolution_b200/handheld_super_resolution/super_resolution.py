"""Alg. 1 (HandheldBurstSuperResolution) — mirrors handheld_super_resolution/super_resolution.py of the reference:
main (:41-200) is the device pipeline, process (:203-360) the host wrapper around it."""
import os
import time

import numpy as np
import torch

from .alignment import align, init_alignment
from .kernels import estimate_kernels
from .merge import fast_path_applies, merge, merge_batch, merge_ref
from .params import sanitize_config, update_snr_config
from .robustness import compute_robustness, init_robustness
from .utils import add_many, timer
from .utils_image import compute_grey_images


_COPY_STREAMS = {}
PHASE_EVENTS = None     # set to a list to collect (phase name, CUDA event) marks of main() (tools/phase_breakdown.py)


def _mark(name):
    if PHASE_EVENTS is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PHASE_EVENTS.append((name, ev))



def _copy_stream(device):
    """One persistent copy stream per device: a fresh torch stream per call would park the staging buffers in a
    different per-stream allocator pool every time and force cudaMalloc in steady state."""
    s = _COPY_STREAMS.get(device.index)
    if s is None:
        s = _COPY_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return s


_INFLIGHT = {}      # device index -> events marking the end of the bursts enqueued so far (most recent last)
MAX_BURSTS_IN_FLIGHT = int(os.environ.get("HHSR_MAX_BURSTS_IN_FLIGHT", "2"))


def _throttle_host(device):
    """main() only enqueues work; a caller that loops over bursts without ever reading a result would queue work without
    bound (the driver then blocks the launching thread until its queues DRAIN, which leaves the GPU idle while they are
    refilled — measured: one 90 ms stall every few bursts).  Entering main() therefore waits until at most
    MAX_BURSTS_IN_FLIGHT earlier bursts are still running on this device: no cost in steady state, the host just stays at
    most that far ahead of the GPU."""
    q = _INFLIGHT.setdefault(device.index, [])
    while len(q) >= max(1, MAX_BURSTS_IN_FLIGHT):
        q.pop(0).synchronize()


MAX_FRAMES_IN_FLIGHT = int(os.environ.get("HHSR_MAX_FRAMES_IN_FLIGHT", "10"))


def _frame_enqueued(device, k):
    """Same idea at frame granularity, inside the frame loop: the host never runs more than MAX_FRAMES_IN_FLIGHT comp
    frames (~95 launches each) ahead of the compute stream, which keeps the driver's launch queues from ever filling up
    (0 switches it off).  The host needs 0.45 ms to enqueue a frame the GPU works 0.95 ms on, so it waits here often — on
    an event, not inside a launch call."""
    if MAX_FRAMES_IN_FLIGHT <= 0:
        return
    q = _INFLIGHT.setdefault(("frames", device.index), [])
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    q.append(ev)
    while len(q) > MAX_FRAMES_IN_FLIGHT:
        q.pop(0).synchronize()


def _burst_enqueued(device):
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    _INFLIGHT.setdefault(device.index, []).append(ev)


_ALIGN_STREAMS = {}


def _align_stream(device, i=0):
    """Persistent high-priority side streams on which the alignment chains of the next frames (grey image, pyramid,
    block matching, ICA: many small latency-bound launches) run while the main stream weights and merges frame k."""
    s = _ALIGN_STREAMS.get((device.index, i))
    if s is None:
        s = _ALIGN_STREAMS[(device.index, i)] = torch.cuda.Stream(device=device, priority=-1)
    return s


ALIGN_AHEAD = int(os.environ.get("HHSR_ALIGN_AHEAD", "1"))   # alignment chains in flight ahead of the merge (0: one stream)
# comp frames merged per pass over the accumulators (merge_batch): their raw / flow / covariance / robustness arrays
# stay resident (144 MB per 12 MP frame) until the batch is merged.  1 = the reference's one launch per frame.
# 0 = automatic: the whole burst in one pass (up to 24 frames) when the frames are already on the device — least
# accumulator traffic — and 5 frames per pass when they stream in from the host, so that merging overlaps the uploads
# and only the last short batch is left when the last frame arrives.
MERGE_BATCH = int(os.environ.get("HHSR_MERGE_BATCH", "0"))
# the last batch of a burst also merges the reference frame and divides (merge.merge_batch(finish=...)): num / den never
# return to HBM.  False keeps merge_ref + divide as a separate pass (the reference's structure; tests compare the two).
FUSE_FINISH = True


def _host_tensor(frame):
    """numpy / torch host frame -> contiguous host tensor (float32, or uint16 sensor counts kept as they are)."""
    if isinstance(frame, torch.Tensor):
        t = frame
    else:
        a = np.ascontiguousarray(frame)
        if a.dtype == np.uint16:
            t = torch.from_numpy(a.view(np.int16)).view(torch.uint16)
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    if t.dtype not in (torch.float32, torch.uint16):
        t = t.to(torch.float32)
    return t if t.is_contiguous() else t.contiguous()


class FrameFeeder:
    """Frames of a burst -> normalised float32 CUDA tensors, streamed.

    Host frames are copied on a persistent copy stream into a small ring of device staging buffers (allocated on the
    compute stream, so the caching allocator serves them from the same pool every burst), one frame ahead of the
    compute stream; uint16 frames (sensor counts) cross PCIe as 2 bytes per pixel and are normalised on the device
    (utils_dng.RawNormalization, the reference's utils_dng.py:146-160).  CUDA float32 frames pass through."""
    # frame being merged + frames being aligned + frame being uploaded (+ batch - 1 frames waiting in a merge batch)
    # The uploads of the next frames (and of the next burst) run ahead of the compute stream as far as free slots allow.
    SLOTS = int(os.environ.get("HHSR_STAGING_SLOTS", "12"))
    _RINGS = {}     # (device, compute stream, role, shape, dtype, slot) -> [buffer, event "slot free"]: staging buffers live
                    # across bursts, so the first uploads of burst i+1 need not wait for the compute stream to drain burst i
    _CURSORS = {}   # (device, compute stream, role, shape, dtype) -> frames staged so far (ring position, kept across bursts)

    def __init__(self, frames, ids, config, device, role="comp", extra_slots=0):
        self.frames, self.ids, self.config, self.device, self.role = frames, list(ids), config, device, role
        # frames waiting in a merge batch keep their slots; the reference frame is cloned at once (two slots: bursts overlap)
        self.SLOTS = 2 if role == "ref" else FrameFeeder.SLOTS + extra_slots
        self.compute = torch.cuda.current_stream(device)
        self.copy = _copy_stream(device)
        self.norm = None
        self.ready = {}

    def _slot(self, host):
        """Next staging buffer of the ring for frames shaped like `host`.  The ring position is kept ACROSS bursts: the
        first frames of burst i + 1 take the buffers released longest ago (early frames of burst i, merged long since)
        instead of the ones its last merge batch still holds, so their uploads overlap the tail of burst i."""
        ring = (self.device.index, self.compute.cuda_stream, self.role, tuple(host.shape), host.dtype)
        pos = FrameFeeder._CURSORS.get(ring, 0)
        FrameFeeder._CURSORS[ring] = pos + 1
        key = ring + (pos % self.SLOTS,)
        slot = FrameFeeder._RINGS.get(key)
        if slot is None:
            # first touch of this ring (or a longer ring than before): allocate EVERY missing slot now.  A fresh block of
            # the compute-stream pool may still be in use by work already queued there, so its first upload has to wait
            # for the compute stream to reach this point — once, not once per burst while the ring fills up
            if len(FrameFeeder._RINGS) > 256:      # shapes changed many times: drop the old staging buffers
                FrameFeeder._RINGS.clear()
            ev = torch.cuda.Event()
            for i in range(self.SLOTS):
                if ring + (i,) not in FrameFeeder._RINGS:
                    # uint16 sensor counts: the slot also owns the float32 frame they are normalised into, so a steady
                    # stream of bursts allocates nothing (a cudaMalloc in the middle of a burst stalls the whole pipeline)
                    norm = torch.empty(host.shape, dtype=torch.float32, device=self.device) if host.dtype == torch.uint16 else None
                    FrameFeeder._RINGS[ring + (i,)] = [torch.empty(host.shape, dtype=host.dtype, device=self.device), ev, norm]
            ev.record(self.compute)
            slot = FrameFeeder._RINGS[key]
        return slot

    def owns(self, t):
        return any(t is s[0] or t is s[2] for s in FrameFeeder._RINGS.values())

    def _stage(self, k):
        """Enqueue the H2D copy of the k-th frame of `ids` (no-op for device frames)."""
        if k >= len(self.ids) or k in self.ready:
            return
        frame = self.frames[self.ids[k]]
        if isinstance(frame, torch.Tensor) and frame.is_cuda:
            self.ready[k] = (frame, None)
            return
        host = _host_tensor(frame)
        slot = self._slot(host)
        buf = slot[0]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(slot[1])                 # previous user of this slot is done
            buf.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        self.ready[k] = (buf, ev, slot)

    def get(self, k):
        """k-th frame as float32 on the compute stream; prefetches frame k + 1.  Call release(k) after its last use."""
        self._stage(k)
        self._stage(k + 1)
        item = self.ready[k]
        if item[1] is None:
            return self._normalised(item[0])
        buf, ev, slot = item
        self.compute.wait_event(ev)
        return self._normalised(buf, slot[2])

    def _normalised(self, t, out=None):
        if t.dtype == torch.uint16:
            if self.norm is None:
                from .utils_dng import RawNormalization
                self.norm = RawNormalization.from_config(self.config)
            return self.norm.apply(t, out=out)
        return t if t.dtype == torch.float32 else t.to(torch.float32)

    def release(self, k):
        item = self.ready.pop(k, None)
        if item is not None and item[1] is not None:
            ev = torch.cuda.Event()
            ev.record(self.compute)
            item[2][1] = ev


def main(ref_img, comp_imgs, config, frame_ids=None, reduce_fn=None, accumulators=None, finalize_fn=None, merge_batch_size=None,
         frame_sink=None, align_ahead=None):
    """Device pipeline (super_resolution.py:41-200).

    ref_img [H,W], comp_imgs [N-1,H,W]: float32 host arrays (numpy / pinned torch) or CUDA tensors.
    Returns (num/den as a CUDA tensor [round(s*H), round(s*W), 3] float32, debug_dict) like the reference.

    B200 additions (optional, used by the distributed driver): `frame_ids` restricts the comp loop to a subset of
    frames (frame sharding) and `reduce_fn(num, den, acc_rob)` is called once after the loop — the one natural
    reduction point of the pipeline.  `accumulators=(num, den)` supplies caller-owned accumulators
    (e.g. peer-mapped symmetric memory; zeroed here) and `finalize_fn(ref_img, covs, num, den, acc_rob, cfa, config)`
    replaces reduction + merge_ref + divide by one fused step (distributed.P2PReduce) and returns the image.
    `merge_batch_size` (default MERGE_BATCH): comp frames accumulated per pass over num/den (merge.merge_batch; the result
    is bit-identical for every batch size).  `frame_sink` (with finalize_fn): an object with `outputs(k) -> (r, covs)`
    (buffers the k-th frame's robustness and covariances are written into) and `publish(k, im_id, frame, flow, covs, r)`:
    the aligned frames are handed over instead of being merged here and no accumulators are allocated — the row-sharded
    multi-GPU merge (distributed.RowShardedMerge) merges all frames of all ranks into this rank's slice of output rows.
    `align_ahead`: alignment chains in flight on side streams (default ALIGN_AHEAD = 1)."""
    verbose_2 = config.verbose >= 2
    grey_method = config.grey_method
    _throttle_host(torch.device("cuda", torch.cuda.current_device()))
    if config.mode != "bayer":
        raise NotImplementedError("only bayer mode is supported (grey mode is broken upstream, SURVEY Q14)")
    init_alignment_ = timer(init_alignment, verbose_2, "\nInitializing alignment", "Alignment initialized (Total)")
    init_robustness_ = timer(init_robustness, verbose_2, "\nEstimating ref image local stats", "Local stats estimated (Total)")
    align_ = timer(align, verbose_2, "\nBeginning alignment", "Image aligned (Total)")
    compute_robustness_ = timer(compute_robustness, verbose_2, "\nEstimating robustness", "Robustness estimated (Total)")
    estimate_kernels_ = timer(estimate_kernels, verbose_2, "\nEstimating kernels", "Kernels estimated (Total)")
    merge_ = timer(merge, verbose_2, "\nAccumulating Image", "Image accumulated (Total)")
    merge_batch_ = timer(merge_batch, verbose_2, "\nAccumulating Image", "Image accumulated (Total)")
    merge_ref_ = timer(merge_ref, verbose_2, "\nAccumulating ref Img", "Ref Img accumulated (Total)")

    debug_mode = config.debug
    debug_dict = {"robustness": [], "flow": []}
    accumulate_r = config.accumulated_robustness_denoiser.enabled or config.robustness.save_mask
    dev = torch.device("cuda", torch.cuda.current_device())
    t1 = time.perf_counter()

    _mark("start")
    ref_feed = FrameFeeder([ref_img], [0], config, dev, role="ref")
    cuda_ref_img = ref_feed.get(0)
    if ref_feed.owns(cuda_ref_img):
        cuda_ref_img = cuda_ref_img.clone()      # the reference frame outlives its staging slot
    ref_feed.release(0)
    cfa_pattern = config.exif.cfa_pattern
    white_balance = config.exif.white_balance
    from .robustness import noise_table
    noise_tab = noise_table((np.asarray(config.noise_model.std_curve, dtype=np.float64),
                             np.asarray(config.noise_model.diff_curve, dtype=np.float64))) if config.robustness.enabled else None

    cuda_ref_grey = compute_grey_images(cuda_ref_img, grey_method)
    ref_pyramid, tyled_pyr, ref_tiled_fft, ref_gradx, ref_grady, ref_hessian = init_alignment_(cuda_ref_grey, config)
    ref_local_means, ref_local_stds = init_robustness_(cuda_ref_img, cfa_pattern, white_balance, config,
                                                       noise_model=noise_tab, need_stds=False)

    H, W = cuda_ref_img.shape
    scale = config.scale
    output_size = (round(scale * H), round(scale * W))
    n_images = len(comp_imgs)
    ids = list(range(n_images) if frame_ids is None else frame_ids)
    # the first comp frame initialises the accumulators (merge(init=True)); only a burst/shard without comp frames
    # needs them zero-filled (the reference uploads host zeros, super_resolution.py:123-124)
    if frame_sink is not None:
        assert finalize_fn is not None, "frame_sink needs a finalize_fn that produces the image"
        num = den = None
    elif accumulators is not None:
        num, den = accumulators
        assert tuple(num.shape) == (*output_size, 3) and tuple(den.shape) == (*output_size, 3)
        if not ids:
            num.zero_(), den.zero_()
    else:
        alloc = torch.empty if ids else torch.zeros
        num = alloc((*output_size, 3), dtype=torch.float32, device=dev)
        den = alloc((*output_size, 3), dtype=torch.float32, device=dev)
    accumulated_r = torch.zeros((H, W), dtype=torch.float64, device=dev) if accumulate_r else None

    _mark("ref_side")
    batch_size = int(MERGE_BATCH if merge_batch_size is None else merge_batch_size)
    if batch_size <= 0:
        on_device = all(isinstance(comp_imgs[i], torch.Tensor) and comp_imgs[i].is_cuda for i in ids)
        batch_size = 24 if on_device else 5
    feed = FrameFeeder(comp_imgs, ids, config, dev, extra_slots=batch_size - 1)
    pending = []        # (k, frame, flow, covs, r) of the frames waiting for the next pass over the accumulators
    merged_any = False
    # the last batch can absorb merge_ref + divide (one pass less over num / den): plain single-GPU runs on the fast path
    fuse_finish = (FUSE_FINISH and frame_sink is None and finalize_fn is None and reduce_fn is None and len(ids) > 0
                   and not config.accumulated_robustness_denoiser.enabled
                   and fast_path_applies(H, W, scale, config.block_matching.tuning.tile_size))
    covs_ref = estimate_kernels_(cuda_ref_img, config) if fuse_finish else None
    main_stream = torch.cuda.current_stream(dev)
    # alignment chains in flight ahead of the merge, each on its own side stream.  One is enough to keep a single GPU busy
    # (more measure the same); a rank of a multi-GPU run has 2-3 frames and latency-bound chains, so it runs them all at once
    ahead = 0 if (os.environ.get("HHSR_SINGLE_STREAM", "0") == "1" or verbose_2) else (ALIGN_AHEAD if align_ahead is None else int(align_ahead))

    def start_alignment(k):
        """Frame k: H2D wait (+ uint16 normalisation) on the main stream, then grey image and alignment on one of the
        alignment streams.  Returns (frame, flow, event marking the flow ready)."""
        cuda_img = feed.get(k)          # H2D of frame k+1 overlaps the work on the earlier frames
        if ahead == 0:
            grey = compute_grey_images(cuda_img, grey_method)
            return cuda_img, align_(ref_pyramid, tyled_pyr, ref_tiled_fft, ref_gradx, ref_grady, ref_hessian, grey, config), None
        align_stream = _align_stream(dev, k % ahead)
        ready = torch.cuda.Event()
        ready.record(main_stream)       # frame (and, the first time, the reference-side products) are ready
        with torch.cuda.stream(align_stream):
            align_stream.wait_event(ready)
            grey = compute_grey_images(cuda_img, grey_method)
            flow = align_(ref_pyramid, tyled_pyr, ref_tiled_fft, ref_gradx, ref_grady, ref_hessian, grey, config)
            done = torch.cuda.Event()
            done.record(align_stream)
        return cuda_img, flow, done

    from collections import deque
    depth = max(ahead, 1)
    inflight = deque(start_alignment(j) for j in range(min(depth, len(ids))))
    for k, im_id in enumerate(ids):
        cuda_img, flow, done = inflight.popleft()
        if k + depth < len(ids):
            inflight.append(start_alignment(k + depth))     # overlaps the rest of this iteration (and the next ones)
        if done is not None:
            main_stream.wait_event(done)
            flow.record_stream(main_stream)
        if debug_mode:
            debug_dict["flow"].append(flow.cpu().numpy())
        r_out, covs_out = frame_sink.outputs(k) if frame_sink is not None else (None, None)
        r = compute_robustness_(cuda_img, ref_local_means, ref_local_stds, flow, cfa_pattern, white_balance,
                                noise_tab, config, out=r_out)
        covs = estimate_kernels_(cuda_img, config, out=covs_out)
        if frame_sink is not None:
            frame_sink.publish(k, im_id, cuda_img, flow, covs, r)
            if accumulate_r:
                add_many(accumulated_r, [r])
            feed.release(k)
            _frame_enqueued(dev, k)
            continue
        pending.append((k, cuda_img, flow, covs, r))
        if len(pending) == batch_size or k == len(ids) - 1:
            # one pass over num/den for the whole batch; the first batch of a burst initialises them
            if fuse_finish and k == len(ids) - 1:
                merge_batch_([p[1] for p in pending], [p[2] for p in pending], [p[3] for p in pending],
                             [p[4] for p in pending], num, den, cfa_pattern, config, init=not merged_any,
                             finish=(cuda_ref_img, covs_ref))
            elif len(pending) == 1:
                merge_(cuda_img, flow, covs, r, num, den, cfa_pattern, config, init=not merged_any)
            else:
                merge_batch_([p[1] for p in pending], [p[2] for p in pending], [p[3] for p in pending],
                             [p[4] for p in pending], num, den, cfa_pattern, config, init=not merged_any)
            merged_any = True
            if accumulate_r:      # accumulated_r += r (super_resolution.py:159), in frame order, one pass per batch
                add_many(accumulated_r, [p[4] for p in pending])
            for p in pending:
                feed.release(p[0])
            pending = []
        if debug_mode:
            debug_dict["robustness"].append(r.cpu().numpy())
        _frame_enqueued(dev, k)

    # the one reduction point of the pipeline (frame-sharded runs): reduce_fn sums the accumulators across ranks and
    # may hand back the slice of output rows this rank has to normalise plus a callable that re-assembles the image
    rows, gather_fn = None, None
    _mark("frames")
    if finalize_fn is not None:
        covs = estimate_kernels_(cuda_ref_img, config)
        num = finalize_fn(cuda_ref_img, covs, num, den, accumulated_r, cfa_pattern, config)
        _mark("fused_reduce_merge_ref")
        if accumulate_r:
            debug_dict["accumulated robustness"] = accumulated_r
        _burst_enqueued(dev)
        return num, debug_dict
    if reduce_fn is not None:
        res = reduce_fn(num, den, accumulated_r)
        if res is not None:
            rows, gather_fn = res
    _mark("reduce")

    if not fuse_finish:
        covs = estimate_kernels_(cuda_ref_img, config)
        use_acc = accumulated_r if config.accumulated_robustness_denoiser.enabled else None
        merge_ref_(cuda_ref_img, covs, num, den, cfa_pattern, config, use_acc, fuse_divide=True, rows=rows)   # + utils.divide, :191
    _mark("merge_ref")
    if gather_fn is not None:
        gather_fn(num)
    _mark("gather")

    if config.verbose >= 1:
        torch.cuda.synchronize()
        s = "\nTotal ellapsed time : "
        print(s, " " * (50 - len(s)), ": ", round((time.perf_counter() - t1), 2), "seconds")
    if accumulate_r:
        debug_dict["accumulated robustness"] = accumulated_r
    _burst_enqueued(dev)
    return num, debug_dict


def load_burst(burst_path):
    """Host loader for the drop-in `process()`: a directory (or .npz) holding a RAW burst as arrays.
    DNG decoding (utils_dng.py) is outside this repository's scope."""
    p = str(burst_path)
    if os.path.isdir(p):
        p = os.path.join(p, "burst.npz")
    if not os.path.exists(p):
        raise FileNotFoundError("expected a burst archive at %s (keys: burst [N,H,W] float32 in [0,1], optional "
                                "cfa_pattern, white_balance, alpha, beta, iso, std_curve, diff_curve)" % p)
    z = np.load(p)
    return {k: z[k] for k in z.files}


def process(burst_path, config, output_dtype=None):
    """Host wrapper (super_resolution.py:203-360): load the burst, derive the SNR-based parameters exactly like the
    reference (mutating `config`), run main(), then — on the DEVICE, where the reference goes through the host — the
    frame-count-aware denoisers (:317-331) and raw2rgb.postprocess (:336-349) when the configuration enables them, and
    return (np.ndarray [H*s, W*s, 3], debug_dict).  RAW decoding is out of scope: the burst comes from an .npz archive
    (see load_burst; optional keys xyz2cam and orientation stand in for the EXIF tags the reference reads).
    output_dtype (B200 addition): None / "float32" returns what the reference's process() returns; "uint8" / "uint16"
    return the quantised image run_handheld.py saves (nan_to_num, clip, rint(x * 255 | 65535)) — a quarter / half of the
    device-to-host bytes."""
    from . import raw2rgb
    from .config import Config
    from .noise_model import run_fast_MC
    from .utils_image import apply_orientation, frame_count_denoising_gauss, frame_count_denoising_median
    data = load_burst(burst_path)
    burst = np.asarray(data["burst"])
    cfa = np.asarray(data.get("cfa_pattern", [[0, 1], [1, 2]])).tolist()
    wb = np.asarray(data.get("white_balance", [1.0, 1.0, 1.0, 0.0]), dtype=np.float64).tolist()
    raw_levels = None
    if np.issubdtype(burst.dtype, np.integer):
        # sensor counts: normalised like utils_dng.py:146-160 — on the device, frame by frame, inside main();
        # only the reference frame is also normalised here, for the SNR estimate below
        from .utils_dng import RawNormalization
        if "black_levels" not in data or "white_level" not in data:
            raise ValueError("integer burst: the archive must provide black_levels and white_level")
        raw_levels = (np.asarray(data["black_levels"]).reshape(-1).tolist(), int(data["white_level"]))
        if burst.dtype != np.uint16:
            # the device path moves uint16 counts; anything else must fit them (the reference casts to float32 without loss)
            lo, hi = int(burst.min()), int(burst.max())
            if lo < 0 or hi > 65535 or raw_levels[1] > 65535:
                raise ValueError("integer burst with samples in [%d, %d] (white level %d) does not fit uint16 sensor counts"
                                 % (lo, hi, raw_levels[1]))
            burst = burst.astype(np.uint16)
        ref_raw = RawNormalization(cfa, raw_levels[0], raw_levels[1], wb).apply_numpy(burst[0])
        ref_in, raw_comp = burst[0], burst[1:]
    else:
        burst = burst.astype(np.float32)
        ref_raw = ref_in = burst[0]
        raw_comp = burst[1:]
    if config.noise_model.get("alpha", None) is None:
        if "alpha" not in data:
            raise ValueError("noise model: alpha/beta neither in the config nor in the burst archive")
        config.noise_model.update({"alpha": float(data["alpha"]), "beta": float(data["beta"])})
    alpha, beta = config.noise_model.alpha, config.noise_model.beta
    if "std_curve" in data:
        std_curve, diff_curve = np.asarray(data["std_curve"]), np.asarray(data["diff_curve"])
    else:
        std_curve, diff_curve = run_fast_MC(alpha, beta)
    brightness = np.mean(ref_raw)
    SNR = brightness / std_curve[round(1000 * brightness)]
    update_snr_config(config, SNR)
    sanitize_config(config, ref_raw.shape)
    config.exif = Config.wrap({"cfa_pattern": cfa, "iso": int(data.get("iso", 100)), "white_balance": wb})
    if raw_levels is not None:
        config.exif.black_levels, config.exif.white_level = raw_levels
    config.noise_model.update({"std_curve": std_curve.tolist(), "diff_curve": diff_curve.tolist()})
    ard = config.accumulated_robustness_denoiser
    ard.enabled = bool(any(x.enabled for x in (ard.median, ard.gauss, ard.merge)))
    out, debug_dict = main(ref_in, raw_comp, config)
    if ard.median.enabled:            # super_resolution.py:324-326
        out = frame_count_denoising_median(out, debug_dict["accumulated robustness"], ard.median, scale=config.scale, mode=config.mode)
    if ard.gauss.enabled:             # :327-329
        out = frame_count_denoising_gauss(out, debug_dict["accumulated robustness"], ard.gauss, scale=config.scale, mode=config.mode)
    post = config.get("postprocessing", None)
    if post is not None and post.enabled:      # :336-349
        out = raw2rgb.postprocess(None, out, post.do_color_correction, post.do_tonemapping, post.do_gamma_correction,
                                  post.sharpening, post.do_devignetting, data.get("xyz2cam", np.zeros((3, 3))),
                                  output_dtype=output_dtype)
    elif output_dtype not in (None, "float32"):
        out = raw2rgb.postprocess(None, out, False, False, False, None, False, None, output_dtype=output_dtype)
    ori = int(data.get("orientation", 1))      # EXIF 'Image Orientation' (:352-360)
    output_image = apply_orientation(out.cpu().numpy(), ori)
    if "accumulated robustness" in debug_dict:
        debug_dict["accumulated robustness"] = apply_orientation(debug_dict["accumulated robustness"].cpu().numpy(), ori)
    return output_image, debug_dict
