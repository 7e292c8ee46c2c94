"""Benchmark of the burst super-resolution hot path (BASELINE.json metric: output MPix/s for a 20-frame 12 MP
Bayer burst merged to 48 MP, scale 2; plus the merge kernel's achieved HBM GB/s against the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repository's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # frame-sharded, one rank per GPU
    python bench.py --impl reference [...]                           # CPU arm: the oracle port on the host cores

A "step" is one whole burst through main(): reference-side products, then every comp frame aligned, weighted
and merged, reference merge and normalisation.  `value` times the step with the burst already resident in HBM;
`e2e` times the same call with the burst in pinned host memory (H2D of every frame and D2H of the 48 MP result
inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # before the CUDA context exists (see handheld_super_resolution/__init__.py)
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {   # BASELINE.json configs
    "20x12MP_s2": dict(n=20, H=3000, W=4000, scale=2),
    "8x12MP_s2": dict(n=8, H=3000, W=4000, scale=2),
    "13x12MP_s3": dict(n=13, H=3000, W=4000, scale=3),
    "20x50MP_s2": dict(n=20, H=6144, W=8192, scale=2),
    "2x256_s1": dict(n=2, H=256, W=256, scale=1),
}


def noise_curves():
    """The reference's own ISO-100 noise curves (data/noise_model_{std,diff}_ISO_100.npy of the reference; SURVEY 8d
    fixes them as the benchmark input) from the committed copy tests/golden/noise_curves_iso100.npz."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "noise_curves_iso100.npz"))
    return z["std_curve"], z["diff_curve"]


def make_config(scale, H, W, ref_brightness):
    """configs/defaults.yaml + the benchmark settings of SURVEY section 8d: tile_size 32 explicit, ISO-100 noise model
    and curves, RGGB, merge constants derived from the SNR exactly as process() derives it
    (super_resolution.py:261-274: brightness = mean of the reference frame, SNR = brightness / std_curve[...])."""
    from handheld_super_resolution.config import Config, load_config
    from handheld_super_resolution.params import sanitize_config, update_snr_config
    from handheld_super_resolution.synthetic import ALPHA_ISO100, BETA_ISO100, CFA_RGGB, WHITE_BALANCE
    cfg = load_config(overrides={"scale": scale, "verbose": 0, "block_matching": {"tuning": {"tile_size": 32}}})
    if min(H, W) < 673:   # default factors need >= 673 px; small plumbing config uses [1,2,2,2]
        cfg.block_matching.tuning.factors = [1, 2, 2, 2]
    cfg.noise_model.alpha, cfg.noise_model.beta = ALPHA_ISO100, BETA_ISO100
    std_curve, diff_curve = noise_curves()
    update_snr_config(cfg, float(ref_brightness) / std_curve[round(1000 * float(ref_brightness))])
    cfg.exif = Config.wrap({"cfa_pattern": CFA_RGGB, "iso": 100, "white_balance": WHITE_BALANCE})
    cfg.noise_model.std_curve, cfg.noise_model.diff_curve = std_curve, diff_curve
    cfg.accumulated_robustness_denoiser.enabled = False
    sanitize_config(cfg, (H, W))
    return cfg


def workload_config(wl_name, wl, n_gpus):
    """The `config` object of the JSON line — the same for both arms (the reference arm runs a bounded sample of it,
    described in its cpu_baseline.sample)."""
    return {"workload": wl_name, **wl, "tile_size": 32, "noise_curves": "reference data/noise_model_*_ISO_100.npy",
            "snr": "derived from the reference frame as process() does",
            "parallelism": "single GPU" if n_gpus == 1 else "comp frames sharded over %d GPUs, one reduction point" % n_gpus,
            "l2": "inputs per step (%.0f MB burst + %.0f MB accumulators) exceed the 126 MB L2; no flush needed"
                  % (wl["n"] * wl["H"] * wl["W"] * 4 / 1e6, round(wl["scale"] * wl["H"]) * round(wl["scale"] * wl["W"]) * 24 / 1e6)}


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML queried in-process
    every 50 ms from a thread (nvidia_ml_py), falling back to one resident `nvidia-smi -lms` process when NVML cannot be
    loaded.  (A resident nvidia-smi polling at 10 Hz was seen to stall the launching thread of short legs at random;
    the in-process queries are three cheap calls per sample.)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip() != ""]
        if ids and all(v.isdigit() for v in ids) and self.index < len(ids):
            return int(ids[self.index])
        return self.index

    def start(self):
        if os.environ.get("HHSR_BENCH_SAMPLER", "nvml") == "off":
            return
        try:
            if os.environ.get("HHSR_BENCH_SAMPLER", "nvml") != "nvml":
                raise RuntimeError("nvidia-smi requested")
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index()))
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml
        R = nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown
        bits = [("hw_slowdown", R),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0))),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0))),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0)))]
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = reasons_fn(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                self.rows.append((time.perf_counter(), [str(sm), str(mx), "%.2f" % pw] + ["Active" if (mask & b) else "Not Active" for _, b in bits]))
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sampler (nvml and nvidia-smi unavailable or switched off)"]}
        time.sleep(0.15)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows),
                "source": "nvml in-process, 50 ms" if self.nvml is not None else "nvidia-smi -lms 100"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def merge_algorithmic_bytes(H, W, scale, ny, nx, K=1, init=False, finish=False):
    """SURVEY section 8d / DESIGN.md, one merge launch over K comp frames: ONE pass over num+den (24 B per HR pixel
    written, and read as well unless the launch initialises them) + per frame raw, r, covariances (4 + 4 + 16/4 = 12 B
    per LR pixel) and the tile flow:  B_batch(K) = HR*24*(1 + [not init]) + K*(LR*12 + tiles*8)  — never K * B_frame.
    A finishing launch (last batch fused with merge_ref + divide) writes the 12 B per HR pixel of the image instead of
    num + den and also reads the reference frame and its covariances (8 B per LR pixel)."""
    hs, ws = round(scale * H), round(scale * W)
    if finish:
        return hs * ws * (12 + (0 if init else 24)) + K * (H * W * 12 + ny * nx * 8) + H * W * 8
    return hs * ws * 24 * (1 if init else 2) + K * (H * W * 12 + ny * nx * 8)


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle (a port of the reference; the reference itself is Numba-CUDA only) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _oracle_frame(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hhsr_oracle as O
    ref, rm, rs, img, cfg = args
    cfa, wb = cfg["exif"]["cfa_pattern"], cfg["exif"]["white_balance"]
    flow = O.align(ref, O.grey_fft(img), cfg)
    r = O.compute_robustness(img, rm, rs, flow, cfa, wb, cfg["noise_model"]["std_curve"], cfg["noise_model"]["diff_curve"], cfg)
    covs = O.estimate_kernels(img, cfg)
    H, W = img.shape
    s = cfg["scale"]
    num = np.zeros((round(s * H), round(s * W), 3), np.float32)
    den = np.zeros_like(num)
    O.accumulate(img, flow, covs, r, num, den, cfa, s, cfg["block_matching"]["tuning"]["tile_size"])
    return num, den


def oracle_step(burst, cfg, pool):
    """One burst through the oracle, comp frames spread over the pool's processes."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hhsr_oracle as O
    cfa, wb = cfg["exif"]["cfa_pattern"], cfg["exif"]["white_balance"]
    ref = O.init_alignment(O.grey_fft(burst[0]), cfg)
    rm, rs = O.init_robustness(burst[0], cfa, wb)
    jobs = [(ref, rm, rs, img, cfg) for img in burst[1:]]
    parts = pool.map(_oracle_frame, jobs) if pool is not None else [_oracle_frame(j) for j in jobs]
    num = sum(p[0] for p in parts)
    den = sum(p[1] for p in parts)
    O.accumulate_ref(burst[0], O.estimate_kernels(burst[0], cfg), num, den, cfa, cfg["scale"])
    return O.divide(num, den)


# BASELINE config 1, measured in the build container (CPU only): the reference's own main() under the Numba simulator
CONFIG1_NOTE = ("the reference has no CPU implementation; its own main() under NUMBA_ENABLE_CUDASIM=1 (BASELINE config 1: 2 frames "
                "256x256, scale 1, Ts 32, factors [1,2,2,2]) took 1910.5 s on 8 host cores (GIL-bound, ~1 core busy) = 3.4e-5 output "
                "MPix/s (baseline/run_reference_cudasim.py; output kept as tests/golden/config1_cudasim.npz and checked against "
                "main() in tests/test_gpu_parity.py) - hence the NumPy oracle port as the timed CPU arm")


def cpu_sample(wl, n_frames, crop, cores):
    """Bounded sample of the workload for the CPU arm: an n_frames x crop x crop burst from the same generator, same
    config.  Returns (burst, plain cfg, scaling) with scaling = (sample out MPix) * (n_frames / workload frames): work
    is proportional to pixels x frames, so value = scaling / seconds is the workload-equivalent output MPix/s."""
    from handheld_super_resolution.config import to_plain
    from handheld_super_resolution.synthetic import synth_burst
    burst, _ = synth_burst(n_frames, crop, crop, seed=0)
    cfg = to_plain(make_config(wl["scale"], crop, crop, float(np.mean(burst[0]))))
    cfg["noise_model"]["std_curve"], cfg["noise_model"]["diff_curve"] = noise_curves()
    out_mpix = round(wl["scale"] * crop) ** 2 / 1e6
    return burst, cfg, out_mpix * n_frames / wl["n"]


def run_reference_arm(args, wl, wl_name):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_frames, crop = max(2, min(wl["n"], cores + 1)), 704       # one comp frame per host core (up to the whole burst)
    burst, cfg, scaling = cpu_sample(wl, n_frames, crop, cores)
    nproc = min(cores, n_frames - 1)
    pool = mp.get_context("fork").Pool(nproc) if nproc > 1 else None
    for _ in range(args.warmup):
        oracle_step(burst, cfg, pool)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(burst, cfg, pool)
    dt = (time.perf_counter() - t0) / args.steps
    if pool is not None:
        pool.close()
    value = scaling / dt
    # beside it, when this box has a GPU and the reference copy travelled: the unmodified reference's own Numba-CUDA run
    # of the WHOLE workload (the reference has no CPU implementation; this is the meaningful "beat this" number)
    ref_gpu = {"unavailable": "no CUDA device, Numba or baseline/_ref on this box"}
    script = os.path.join(ROOT, "baseline", "time_reference_numba.py")
    if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "handheld_super_resolution")):
        try:
            import torch
            if torch.cuda.is_available():
                p = subprocess.run([sys.executable, script, str(wl["n"]), str(wl["H"]), str(wl["W"]), str(wl["scale"])],
                                   capture_output=True, text=True, timeout=600)
                js = [x for x in p.stdout.splitlines() if x.startswith("{")]
                ref_gpu = json.loads(js[-1]) if js else {"unavailable": (p.stderr or p.stdout)[-300:]}
        except Exception as e:       # measurement extra: never fails the arm
            ref_gpu = {"unavailable": repr(e)[:300]}
    sample = ("%d-frame %dx%d crop of the workload per step (same generator and config), comp frames spread over %d "
              "processes; value = sample output MPix x (%d/%d frames) / seconds" % (n_frames, crop, crop, nproc, n_frames, wl["n"]))
    line = {"impl": "reference", "metric": "output MPix/s (20x12MP->48MP burst)", "value": value, "unit": "MPix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64 mixed (as the reference)",
            "data": "synthetic", "config": workload_config(wl_name, wl, args.gpus),
            "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": nproc, "kind": "port", "sample": sample, "note": CONFIG1_NOTE},
            "reference_numba_gpu": ref_gpu,
            "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
def host_image_buffers(hs, ws, world, rank, n_buffers=2):
    """Pinned host buffers [hs, ws, 3] float32 for the result.  With several ranks they are views of ONE POSIX
    shared-memory segment mapped by every rank and page-locked with cudaHostRegister: each rank copies its slice of the
    image over its own PCIe link and the consumer (rank 0) sees the whole image.  Falls back to private pinned buffers."""
    import torch
    nbytes = hs * ws * 3 * 4
    if world == 1:
        return [torch.empty((hs, ws, 3), dtype=torch.float32).pin_memory() for _ in range(n_buffers)], "pinned"
    import torch.distributed as dist
    path = "/dev/shm/hhsr_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), hs * ws)
    try:
        if rank == 0:
            with open(path, "wb") as f:
                f.truncate(nbytes * n_buffers)
        dist.barrier()
        flat = torch.from_file(path, shared=True, size=hs * ws * 3 * n_buffers, dtype=torch.float32)
        rc = torch.cuda.cudart().cudaHostRegister(flat.data_ptr(), nbytes * n_buffers, 0)
        assert int(rc) == 0, "cudaHostRegister failed (%s)" % rc
        dist.barrier()
        if rank == 0:
            os.unlink(path)          # the mappings keep the segment alive
        return [flat[i * hs * ws * 3:(i + 1) * hs * ws * 3].view(hs, ws, 3) for i in range(n_buffers)], "shared POSIX shm + cudaHostRegister"
    except Exception as e:
        print("shared host image unavailable (%s); private pinned buffers" % e, file=sys.stderr)
        return [torch.empty((hs, ws, 3), dtype=torch.float32).pin_memory() for _ in range(n_buffers)], "private pinned"


def run_cuda_arm(args, wl, wl_name):
    import torch
    import torch.distributed as dist
    from handheld_super_resolution import _lib, super_resolution as SR
    from handheld_super_resolution.distributed import main_sharded
    from handheld_super_resolution.synthetic import synth_burst

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, H, W, scale = wl["n"], wl["H"], wl["W"], wl["scale"]
    hs, ws = round(scale * H), round(scale * W)
    # the one exchange point of a multi-GPU run: "rows" (merge sharded by output rows over NVLink peer memory, default
    # when symmetric memory can be set up), "p2p" (frame-sharded accumulators summed by one fused peer-memory kernel),
    # "reduce_scatter" / "allreduce" (the same sum over NCCL)
    reduce_mode = args.reduce
    if world > 1 and reduce_mode in ("auto", "rows", "p2p"):
        try:
            from handheld_super_resolution.distributed import P2PReduce, RowShardedMerge
            if reduce_mode in ("auto", "rows"):
                RowShardedMerge.get(H, W, scale, n - 1, 32)
                reduce_mode = "rows"
            else:
                P2PReduce.get((hs, ws, 3))
        except Exception as e:      # no peer access / symmetric memory on this box
            if args.reduce != "auto":
                raise
            print("peer-memory exchange unavailable (%s); using NCCL reduce-scatter" % e, file=sys.stderr)
            reduce_mode = "reduce_scatter"
    elif reduce_mode == "auto":
        reduce_mode = "reduce_scatter"
    os.environ["HHSR_SHARD_REDUCE"] = reduce_mode

    burst_dev, _ = synth_burst(n, H, W, seed=0, device="cuda", as_numpy=False)      # same burst on every rank
    cfg = make_config(scale, H, W, burst_dev[0].mean().item())
    burst_host = torch.empty((n, H, W), dtype=torch.float32).pin_memory()
    burst_host.copy_(burst_dev)
    # the same burst as 14-bit sensor counts (black level 1024): the form a DNG decoder hands over
    import copy
    cfg_u16 = copy.deepcopy(cfg)
    cfg_u16.exif.black_levels, cfg_u16.exif.white_level = [1024, 1024, 1024, 1024], 16383
    wbn = torch.tensor([[cfg.exif.white_balance[c] / cfg.exif.white_balance[1] for c in row] for row in cfg.exif.cfa_pattern],
                       device="cuda", dtype=torch.float32).repeat(H // 2, W // 2)
    counts = torch.round(burst_dev / wbn * (16383 - 1024) + 1024).clamp_(0, 65535).to(torch.int32)
    burst_u16 = torch.empty((n, H, W), dtype=torch.uint16).pin_memory()
    burst_u16.view(torch.int16).copy_(counts.to(torch.int16))   # bit pattern of the low 16 bits
    del counts, wbn
    out_hosts, shared_note = host_image_buffers(hs, ws, world, rank)      # double-buffered D2H target, one image for all ranks
    d2h_stream = torch.cuda.Stream()
    d2h_state = {"k": 0, "events": [None, None]}
    torch.cuda.synchronize()

    # per-launch timing of the dominant kernel (merge accumulate) with CUDA events on the launching stream:
    # (frames in the launch, initialising?, start, end) for every merge launch of the timed steps
    merge_events = []
    orig_merge, orig_merge_batch = SR.merge, SR.merge_batch

    def timed_merge(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_merge(*a, **k)
        e1.record()
        merge_events.append((1, bool(k.get("init")), False, e0, e1))

    def timed_merge_batch(comps, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_merge_batch(comps, *a, **k)
        e1.record()
        merge_events.append((len(comps), bool(k.get("init")), k.get("finish") is not None, e0, e1))
    SR.merge, SR.merge_batch = timed_merge, timed_merge_batch
    batch = args.merge_batch if args.merge_batch > 0 else SR.MERGE_BATCH     # 0: automatic (main() decides per call)
    SR.MERGE_BATCH = batch

    def step_resident():
        out, _ = main_sharded(burst_dev[0], burst_dev[1:], cfg)
        return out

    out8_hosts = [torch.empty((hs, ws, 3), dtype=torch.uint8).pin_memory() for _ in range(2)] if world == 1 else None
    post_cfg = cfg.postprocessing

    def step_e2e_post(pipelined=True):
        """What process() does on this path end to end (SURVEY 8f ranks 1 + 2): uint16 sensor counts in (480 MB),
        main(), the reference's default post-process (unsharp mask + gamma) and uint8 quantisation ON THE DEVICE, 144 MB
        out instead of 576 MB of float32."""
        from handheld_super_resolution import raw2rgb
        out, _ = SR.main(burst_u16[0], burst_u16[1:], cfg_u16)
        img8 = raw2rgb.postprocess(None, out, post_cfg.do_color_correction, post_cfg.do_tonemapping, post_cfg.do_gamma_correction,
                                   post_cfg.sharpening, post_cfg.do_devignetting, None, output_dtype="uint8")
        k = d2h_state["k"] % 2
        d2h_state["k"] += 1
        if d2h_state["events"][k] is not None:
            d2h_state["events"][k].synchronize()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(d2h_stream if pipelined else torch.cuda.current_stream()):
            if pipelined:
                d2h_stream.wait_event(ready)
            out8_hosts[k].copy_(img8, non_blocking=True)
            img8.record_stream(torch.cuda.current_stream())
            done = torch.cuda.Event()
            done.record()
        d2h_state["events"][k] = done
        if not pipelined:
            done.synchronize()

    trace = []      # HHSR_BENCH_TRACE: (ms inside main(), ms waiting for the host buffer, cudaMalloc calls so far) per e2e step

    def step_e2e(pipelined=True, u16=False):
        """Host burst in, host image out.  The D2H of the 48 MP result runs on its own stream into one of two pinned
        buffers, so it overlaps the NEXT burst's compute (steady-state throughput of back-to-back bursts); every
        copy still happens inside the timed region, which ends with a full device synchronisation."""
        tm0 = time.perf_counter()
        if u16:
            out, dbg = main_sharded(burst_u16[0], burst_u16[1:], cfg_u16)
        else:
            out, dbg = main_sharded(burst_host[0], burst_host[1:], cfg)
        tm1 = time.perf_counter()
        # who copies what to the host: with the row-sharded merge every rank owns a slice of the image and sends it over
        # its own PCIe link into the shared host image; otherwise rank 0 holds the whole image
        rows = dbg.get("rows") if isinstance(dbg, dict) else None
        if rows is not None or rank == 0:
            k = d2h_state["k"] % 2
            d2h_state["k"] += 1
            if d2h_state["events"][k] is not None:
                d2h_state["events"][k].synchronize()          # buffer k free again (its previous copy finished)
            if os.environ.get("HHSR_BENCH_TRACE"):
                trace.append((round((tm1 - tm0) * 1e3, 1), round((time.perf_counter() - tm1) * 1e3, 1),
                              torch.cuda.memory_stats().get("num_device_alloc", 0)))
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(d2h_stream if pipelined else torch.cuda.current_stream()):
                if pipelined:
                    d2h_stream.wait_event(ready)
                dst = out_hosts[k] if rows is None else out_hosts[k][rows[0]:rows[1]]
                if out is not None and dst.numel() > 0:
                    dst.copy_(out, non_blocking=True)
                    out.record_stream(torch.cuda.current_stream())
                done = torch.cuda.Event()
                done.record()
            d2h_state["events"][k] = done
            if not pipelined:
                done.synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        stamps = [time.perf_counter()]
        for _ in range(steps):
            fn()
            stamps.append(time.perf_counter())
        if os.environ.get("HHSR_BENCH_TRACE"):       # host-side enqueue time of every step of this leg
            if trace:
                print("   (main ms, buffer-wait ms, mallocs):", trace[-steps:], file=sys.stderr)
                del trace[:]
            print("host ms per step:", [round((b - a) * 1e3, 1) for a, b in zip(stamps, stamps[1:])],
                  "cudaMalloc calls:", torch.cuda.memory_stats().get("num_device_alloc", 0), file=sys.stderr)
        torch.cuda.current_stream().wait_stream(d2h_stream)      # the last result copy belongs to the timed region
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, t0, t1

    def warm(fn):
        """W untimed steps (at least 3), then up to 8 more until a step passes without the caching allocator calling
        cudaMalloc: a leg is timed in its steady state (an allocation in the middle of a burst drains the pipeline for
        ~100 ms; the pools stop growing after a few bursts)."""
        for _ in range(max(args.warmup, 3)):
            fn()
        for _ in range(8):
            before = torch.cuda.memory_stats().get("num_device_alloc", 0)
            fn()
            grew = torch.tensor([float(torch.cuda.memory_stats().get("num_device_alloc", 0) != before)], device="cuda")
            if world > 1:       # every rank must run the same number of steps (they contain the exchange)
                dist.all_reduce(grew, op=dist.ReduceOp.MAX)
            if grew.item() == 0.0:
                break

    warm(step_resident)
    merge_events.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count
    ms_res, t0, t1 = timed(step_resident, args.steps)
    launches = (_lib.launch_count - launches0)
    merge_launches = [(K, init, fin, a.elapsed_time(b)) for K, init, fin, a, b in merge_events]
    merge_events.clear()
    # the reference's launch granularity (one comp frame per pass over the accumulators) for comparison: a few steps
    # with the batching switched off; its read-modify-write launches are the kernel round 1 reported
    per_frame_ms = []
    if batch != 1 and world == 1 and not args.no_e2e:
        SR.MERGE_BATCH = 1
        step_resident()
        merge_events.clear()
        timed(step_resident, 2)
        per_frame_ms = [a.elapsed_time(b) for K, init, fin, a, b in merge_events if not init and not fin]
        merge_events.clear()
        SR.MERGE_BATCH = batch
    if args.no_e2e:        # profiling runs (ncu replays every launch): the resident step only
        ms_e2e = ms_lat = ms_u16 = float("nan")
        ms_post = ms_post_lat = None
        t2 = t1
    else:
        warm(step_e2e)
        ms_e2e, _, t2 = timed(step_e2e, args.steps)
        ms_lat, _, t2 = timed(lambda: step_e2e(pipelined=False), max(2, args.steps // 2))   # one burst at a time, host-synchronous
        warm(lambda: step_e2e(u16=True))
        ms_u16, _, t2 = timed(lambda: step_e2e(u16=True), args.steps)
        ms_post = ms_post_lat = None
        if world == 1:
            warm(step_e2e_post)
            ms_post, _, t2 = timed(step_e2e_post, args.steps)
            ms_post_lat, _, t2 = timed(lambda: step_e2e_post(pipelined=False), max(2, args.steps // 2))
    clocks = sampler.stop(t0, t2) if rank == 0 else None
    SR.merge, SR.merge_batch = orig_merge, orig_merge_batch
    alloc = torch.cuda.memory_stats()
    alloc_note = {"cudaMalloc_calls_total": int(alloc.get("num_device_alloc", 0)), "alloc_retries": int(alloc.get("num_alloc_retries", 0)),
                  "reserved_GB": alloc.get("reserved_bytes.all.peak", 0) / 1e9}

    # correctness of the sharded run, outside every timed region: the same burst through the single-GPU path on this
    # rank, compared with what this rank holds of the sharded result (its row slice, or the whole image on rank 0)
    parity = None
    if world > 1:
        out_s, dbg_s = main_sharded(burst_dev[0], burst_dev[1:], cfg)
        want, _ = SR.main(burst_dev[0], burst_dev[1:], cfg)
        rows = dbg_s.get("rows")
        if rows is not None:
            want = want[rows[0]:rows[1]]
        d = torch.zeros(2, device="cuda")
        if out_s is not None and (rows is not None or rank == 0 or reduce_mode in ("reduce_scatter", "allreduce")) and want.numel() > 0:
            same_nan = torch.equal(torch.isnan(out_s), torch.isnan(want))
            d[0] = torch.nan_to_num((out_s - want).abs(), nan=0.0).max()
            d[1] = 0.0 if same_nan else 1.0
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        parity = {"max_abs_diff": d[0].item(), "nan_pattern_differs": bool(d[1].item() > 0),
                  "note": "sharded result vs the single-GPU path on the same burst, max over ranks; computed outside the timed regions"}
        del out_s, want

    # other BASELINE configurations, resident timing only (driver-observed at 8 GPUs: configs 4 and 5)
    extras = []
    names = args.extra_workloads
    names = (["13x12MP_s3", "20x50MP_s2"] if world >= 8 else []) if names == "auto" else [x for x in names.split(",") if x and x != "none"]
    if names:
        del burst_dev, burst_host, burst_u16      # (symmetric-memory buffers of the main workload stay mapped: a few GB)
        torch.cuda.empty_cache()
    for name in names:
        w2 = WORKLOADS[name]
        try:
            b2, _ = synth_burst(w2["n"], w2["H"], w2["W"], seed=0, device="cuda", as_numpy=False)
            c2 = make_config(w2["scale"], w2["H"], w2["W"], b2[0].mean().item())
            fn = lambda: main_sharded(b2[0], b2[1:], c2)      # noqa: E731
            for _ in range(3):
                fn()
            ms2, _, _ = timed(fn, max(3, args.steps // 2))
            mp = round(w2["scale"] * w2["H"]) * round(w2["scale"] * w2["W"]) / 1e6
            extras.append({"workload": name, **w2, "ms_per_step": ms2, "value": mp / (ms2 * 1e-3), "unit": "MPix/s", "n_gpus": world,
                           "note": "resident burst, same call and exchange mode as the main workload"})
            del b2
            torch.cuda.empty_cache()
        except Exception as e:      # an extra must never cost the main line
            extras.append({"workload": name, "error": repr(e)[:300]})
            break

    if rank == 0:
        out_mpix = hs * ws / 1e6
        peak, peak_src = measured_peak_gbs()
        ny, nx = -(-H // 32), -(-W // 32)
        # roofline of the merge launches of the timed steps: algorithmic bytes B_batch(K) of every launch / its duration
        alg_total = sum(merge_algorithmic_bytes(H, W, scale, ny, nx, K, init, fin) for K, init, fin, _ in merge_launches)
        ms_total = sum(ms for _, _, _, ms in merge_launches)
        frames_total = sum(K for K, _, _, _ in merge_launches)
        achieved = alg_total / (ms_total * 1e-3) / 1e9 if ms_total > 0 else float("nan")
        n_launch = max(len(merge_launches), 1)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "merge_traffic_bytes.json")
        if os.path.exists(tp):
            kmax = max(K for K, _, _, _ in merge_launches) if merge_launches else 0
            fused = any(fin for K, _, fin, _ in merge_launches if K == kmax)      # the finishing launch (merge_ref + divide fused)
            tab = json.load(open(tp))
            traffic = tab.get("%s_batch%d%s" % (wl_name, kmax, "_finish" if fused else "")) or tab.get("%s_batch%d" % (wl_name, kmax))
        alg1 = merge_algorithmic_bytes(H, W, scale, ny, nx)
        pf_ms = float(np.mean(per_frame_ms)) if per_frame_ms else None
        line = {
            "metric": "output MPix/s (20x12MP->48MP burst)" if wl_name == "20x12MP_s2" else "output MPix/s",
            "value": out_mpix / (ms_res * 1e-3), "unit": "MPix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (f64 sub-pixel positions)", "data": "synthetic",
            "config": workload_config(wl_name, wl, world),
            "reduce": None if world == 1 else {
                "rows": "merge sharded by output rows: one exchange of LR row bands over NVLink peer memory, every rank keeps "
                        "its slice of the image (host copies over the ranks' own PCIe links into one %s image)" % shared_note,
                "p2p": "frame-sharded accumulators, one fused peer-memory kernel (NVLink pull + sum + merge_ref + divide), image on rank 0",
                "reduce_scatter": "frame-sharded accumulators, one NCCL reduce-scatter + all-gather of the image",
                "allreduce": "frame-sharded accumulators, one NCCL all-reduce"}[reduce_mode],
            "merge_batch": batch if batch > 0 else "auto (whole burst per pass when resident, 5 frames per pass when streamed from the host)",
            "e2e": {"value": out_mpix / (ms_e2e * 1e-3), "unit": "MPix/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(n * H * W * 4), "d2h_bytes_per_step": int(hs * ws * 3 * 4),
                    "mode": "back-to-back bursts, result D2H double-buffered on a copy stream (overlaps the next burst)",
                    "single_burst_latency_ms": ms_lat, "single_burst_value": out_mpix / (ms_lat * 1e-3),
                    "uint16_raw": {"value": out_mpix / (ms_u16 * 1e-3), "ms_per_step": ms_u16,
                                   "h2d_bytes_per_step": int(n * H * W * 2),
                                   "note": "same call fed with 14-bit sensor counts (uint16), normalised on the device"},
                    "uint16_in_uint8_out": None if ms_post is None else {
                        "value": out_mpix / (ms_post * 1e-3), "ms_per_step": ms_post, "single_burst_latency_ms": ms_post_lat,
                        "h2d_bytes_per_step": int(n * H * W * 2), "d2h_bytes_per_step": int(hs * ws * 3),
                        "note": "the process() path: uint16 sensor counts in, main(), the reference's default post-process (unsharp "
                                "mask + gamma, raw2rgb.py:212-250) and uint8 quantisation on the device, uint8 image out"}},
            "gpu_launches": int(launches),
            "roofline": {"kernel": ("accumulate_pow2_batch_kernel (merge, up to %d comp frames per pass over the accumulators)"
                                    % max(K for K, _, _, _ in merge_launches)) if frames_total > len(merge_launches)
                                   else "accumulate_pow2_kernel (merge, one comp frame per launch)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_total / n_launch, "avg_launch_ms": ms_total / n_launch,
                         "launches_timed": len(merge_launches), "frames_per_launch": frames_total / n_launch,
                         "ms_per_frame": ms_total / max(frames_total, 1),
                         "accounting": "B_batch(K) = HR*24*(1 + [not init]) + K*(LR*12 + tiles*8) per launch (SURVEY 8d; the launch that "
                                       "also merges the reference frame and divides writes HR*12 instead of HR*24 and reads LR*8 more), "
                                       "summed over every merge launch of the timed steps / summed CUDA-event durations; ms_per_frame "
                                       "counts the reference frame's merge + divide inside the finishing launch as part of its frames",
                         "issue_slots_busy_pct_ncu": 78.6,
                         "note": "frame batching removes accumulator traffic: the kernel moves from the HBM roofline to the "
                                 "issue limit of the tap arithmetic, so the step gets faster while this fraction falls",
                         "per_frame_kernel": None if pf_ms is None else {
                             "kernel": "accumulate_pow2_kernel (merge_batch_size=1: one frame per pass, the reference's granularity)",
                             "achieved": alg1 / (pf_ms * 1e-3) / 1e9, "frac": alg1 / (pf_ms * 1e-3) / 1e9 / peak,
                             "avg_launch_ms": pf_ms, "algorithmic_bytes_per_launch": alg1, "launches_timed": len(per_frame_ms)}},
            "clocks": clocks,
            "allocator": alloc_note,
        }
        if parity is not None:
            line["parity_vs_single"] = parity
        if extras:
            line["extra_workloads"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(wl):
    """The oracle port timed on this box's host cores on a bounded sample (single process: `cores` = 1)."""
    burst, cfg, scaling = cpu_sample(wl, 3, 704, 1)
    t0 = time.perf_counter()
    oracle_step(burst, cfg, None)
    dt = time.perf_counter() - t0
    return {"value": scaling / dt, "unit": "MPix/s", "cores": 1, "kind": "port", "seconds": dt, "note": CONFIG1_NOTE,
            "sample": "3-frame 704x704 crop of the workload, NumPy oracle in one process; value = sample output MPix x "
                      "(3/%d frames) / seconds (work ~ pixels x frames)" % wl["n"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="20x12MP_s2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merge-batch", type=int, default=0, help="comp frames per pass over the accumulators (0: the package default)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer legs (e2e fields are NaN)")
    ap.add_argument("--reduce", default="auto", choices=["auto", "rows", "p2p", "reduce_scatter", "allreduce"],
                    help="N > 1: the exchange point (auto = rows, the row-sharded merge, when peer memory is available; "
                         "p2p / reduce_scatter / allreduce sum frame-sharded accumulators)")
    ap.add_argument("--extra-workloads", default="auto", help="comma-separated workloads also timed (resident) at N > 1; "
                    "auto = 13x12MP_s3,20x50MP_s2 at 8 GPUs, none otherwise")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
    else:
        run_cuda_arm(args, wl, args.workload)


if __name__ == "__main__":
    main()
