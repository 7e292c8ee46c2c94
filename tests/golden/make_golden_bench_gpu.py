"""Golden vectors AT THE BENCHMARK SHAPES: runs the UNMODIFIED reference (baseline/_ref, Numba-CUDA + torch) on a
B200 on the bursts bench.py times (3000x4000 scale 2 and 3, 6144x8192 alignment) plus whole pipelines at tile sizes
16 and 64, and stores COMPACT fixtures (< 5 MB in total):

  * the flow of EVERY tile after block matching and after ICA at every pyramid level and for every comp frame
    (block-matching offsets are compared bit-exactly on all tiles, tests/test_gpu_bench_shapes.py);
  * 3x3 grids of crops (corners, edges, centre) of the grey image, robustness r, covariances, accumulators and the
    output image;
  * float64 sums and NaN / zero counts of the full arrays.

The bursts themselves are not stored (48-200 MB per frame): they come from the seeded generator
handheld_super_resolution/synthetic.py on the GPU; their float64 sums and a crop are stored so that the test can prove
it regenerated the same burst bit for bit.

Test tooling, not product code.  Usage (GPU box, through gpurun):

    gpurun -- python tests/golden/make_golden_bench_gpu.py        # writes gpurun_out/golden_bench/*.npz
    cp gpurun_out/golden_bench/*.npz tests/golden/
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
OUT = os.path.join(ROOT, "gpurun_out", "golden_bench")
os.makedirs(OUT, exist_ok=True)

spec = importlib.util.spec_from_file_location("make_golden_gpu", os.path.join(HERE, "make_golden_gpu.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)          # installs the module stubs and puts baseline/_ref on sys.path
to_np = G.to_np
sys.path.insert(0, HERE)
from crop_grid import crop_origins  # noqa: E402


def grid_crops(a, size):
    """-> [9,size,size,...]; the test recomputes the origins with the same crop_origins()."""
    return np.stack([a[y:y + size, x:x + size] for y, x in crop_origins(a.shape, size)]).copy()


def summary(a):
    """[float64 sum over finite entries, NaN count, exact-zero count, inf count]."""
    a64 = a.astype(np.float64)
    fin = np.isfinite(a64)
    return np.array([a64[fin].sum(), np.isnan(a64).sum(), (a64 == 0).sum(), np.isinf(a64).sum()], np.float64)


class Keeper:
    def __init__(self, crop_lr=48, crop_cov=24, crop_hr=32, keep_first=True):
        self.d = {}
        self.sizes = dict(lr=crop_lr, cov=crop_cov, hr=crop_hr)
        self.keep_first = keep_first

    def full(self, name, a):
        self.d[name] = to_np(a)

    def reduced(self, name, a, kind):
        a = to_np(a)
        self.d[name + "__crops"] = grid_crops(a, self.sizes[kind])
        self.d[name + "__sum"] = summary(a)


def run_reference(burst, cfg, K, align_only=False):
    """Reference main() (or only its alignment part) with capture hooks that reduce on the fly."""
    from numba import cuda
    from handheld_super_resolution import super_resolution as SR
    from handheld_super_resolution import alignment as AL
    state = {"frame": 0, "grey": 0, "kern": 0}
    saved = {}

    def patch(mod, name, fn):
        saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, fn)

    o_grey, o_ia, o_al = SR.compute_grey_images, SR.init_alignment, SR.align
    o_l2, o_l1, o_ica = AL.align_lvl_block_matching_L2, AL.align_lvl_block_matching_L1, AL.align_lvl_ica
    o_cr, o_ek, o_m, o_mr = SR.compute_robustness, SR.estimate_kernels, SR.merge, SR.merge_ref
    n_comp = len(burst) - 1

    def grey(img, method):
        out = o_grey(img, method)
        cuda.synchronize()
        if method == "FFT":
            if state["grey"] <= 1:
                K.reduced("grey_%d" % state["grey"], out, "lr")
            state["grey"] += 1
        return out

    def init_al(ref_grey, config):
        out = o_ia(ref_grey, config)
        cuda.synchronize()
        n = len(out[0])
        for i in range(n):  # coarse -> fine
            if i < n - 1:
                K.full("ref_hessian_c%d" % i, out[5][i])
            else:
                h = to_np(out[5][i])
                K.d["ref_hessian_c%d__sum" % i] = summary(h)
                K.d["ref_hessian_c%d__crops" % i] = grid_crops(h, 16)
        return out

    def align(*a, **k):
        state["frame"] += 1
        out = o_al(*a, **k)
        cuda.synchronize()
        return out

    def l2(tyled, fft, moving, alignment, l, config):
        o_l2(tyled, fft, moving, alignment, l, config)
        cuda.synchronize()
        K.full("flow_f%d_l%d_bm" % (state["frame"], l), alignment)

    def l1(ref_lvl, moving, alignments, l, config):
        before = to_np(alignments)
        o_l1(ref_lvl, moving, alignments, l, config)
        cuda.synchronize()
        after = to_np(alignments)
        # for tile sizes 32 / 64 the level is rint(flow); store only whether that held (and the flow if not)
        is_rint = bool(np.array_equal(after, np.rint(before)))
        K.d["flow_f%d_l%d_bm_is_rint" % (state["frame"], l)] = np.array(is_rint)
        if not is_rint:
            K.full("flow_f%d_l%d_in" % (state["frame"], l), before)
            K.full("flow_f%d_l%d_bm" % (state["frame"], l), after)

    def ica(ref_img, gx, gy, hess, moving, alignment, l, config):
        o_ica(ref_img, gx, gy, hess, moving, alignment, l, config)
        cuda.synchronize()
        K.full("flow_f%d_l%d_ica" % (state["frame"], l), alignment)

    def rob(*a, **k):
        out = o_cr(*a, **k)
        cuda.synchronize()
        K.reduced("r_f%d" % state["frame"], out, "lr")
        return out

    def kern(img, config):
        out = o_ek(img, config)
        cuda.synchronize()
        state["kern"] += 1
        K.reduced("covs_%d" % state["kern"], out, "cov")   # 1..N-1 comp frames in order, last = ref
        return out

    def merge(comp, al, covs, r, num, den, cfa, config):
        o_m(comp, al, covs, r, num, den, cfa, config)
        cuda.synchronize()
        # accumulators after the first (main case only) and after the last comp frame
        if state["frame"] == 1 and K.keep_first and n_comp > 1:
            K.reduced("num_first", num, "hr")
            K.reduced("den_first", den, "hr")
        if state["frame"] == n_comp:
            K.reduced("num_comp", num, "hr")
            K.reduced("den_comp", den, "hr")

    def merge_ref(ref, kernels, num, den, cfa, config, acc_rob=None):
        if acc_rob is not None:
            o_mr(ref, kernels, num, den, cfa, config, acc_rob)
        else:
            o_mr(ref, kernels, num, den, cfa, config)
        cuda.synchronize()
        K.reduced("num_final", num, "hr")
        K.reduced("den_final", den, "hr")

    for mod, name, fn in [(SR, "compute_grey_images", grey), (SR, "init_alignment", init_al), (SR, "align", align),
                          (AL, "align_lvl_block_matching_L2", l2), (AL, "align_lvl_block_matching_L1", l1),
                          (AL, "align_lvl_ica", ica), (SR, "compute_robustness", rob),
                          (SR, "estimate_kernels", kern), (SR, "merge", merge), (SR, "merge_ref", merge_ref)]:
        patch(mod, name, fn)
    try:
        if align_only:
            ref_grey = SR.compute_grey_images(cuda.to_device(burst[0]), cfg.grey_method)
            ref = SR.init_alignment(ref_grey, cfg)
            for img in burst[1:]:
                g = SR.compute_grey_images(img, cfg.grey_method)      # host frame, as main() passes it
                SR.align(*ref, g, cfg)
        else:
            out, dbg = SR.main(burst[0], burst[1:], cfg)
            cuda.synchronize()
            K.reduced("out", out, "hr")
            if "accumulated robustness" in dbg:
                K.reduced("acc_rob", dbg["accumulated robustness"], "lr")
    finally:
        for (mod, name), fn in saved.items():
            setattr(mod, name, fn)


def burst_fingerprint(burst):
    return {"burst__sum": np.array([b.astype(np.float64).sum() for b in burst]),
            "burst__crop": burst[-1][100:132, 200:232].copy()}


CASES = [   # name, n, H, W, seed, scale, Ts, metrics override, align_only
    ("bench12_s2", 5, 3000, 4000, 0, 2, 32, None, False),
    ("bench12_s3", 3, 3000, 4000, 0, 3, 32, None, False),
    ("bench50_align", 2, 6144, 8192, 0, 2, 32, None, True),
    ("ts64_pipeline", 3, 2048, 2560, 3, 2, 64, None, False),
    ("ts16_pipeline", 3, 1000, 1400, 4, 2, 16, ["L2", "L2", "L2", "L2"], False),
    ("ts16_l1_probe", 2, 1000, 1400, 4, 2, 16, None, True),      # what the data-racy Ts-16 L1 level returns
]


def main():
    G.install_copysign_shim()
    synth = G.load_synth()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    std_curve = np.load(os.path.join(G.REF, "data", "noise_model_std_ISO_100.npy"))
    diff_curve = np.load(os.path.join(G.REF, "data", "noise_model_diff_ISO_100.npy"))
    report = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    only = sys.argv[1:]
    for name, n, H, W, seed, scale, Ts, metrics, align_only in CASES:
        if only and name not in only:
            continue
        t0 = time.perf_counter()
        burst_t, shifts = synth.synth_burst(n, H, W, seed=seed, device="cuda", as_numpy=False)
        burst = burst_t.cpu().numpy()
        del burst_t
        torch.cuda.empty_cache()
        cfg = G.make_config(scale, Ts, [1, 2, 4, 4], burst[0], std_curve, diff_curve)
        if metrics is not None:
            cfg.block_matching.tuning.metrics = list(metrics)
        K = Keeper() if name == "bench12_s2" else Keeper(crop_lr=32, crop_cov=16, crop_hr=24, keep_first=False)
        K.d.update(burst_fingerprint(burst))
        K.d["shifts"] = np.array(shifts)
        K.d["cfg_json"] = np.array(json.dumps(G.cfg_summary(cfg)))
        K.d["case"] = np.array(json.dumps(dict(n=n, H=H, W=W, seed=seed, scale=scale, Ts=Ts, align_only=align_only)))
        t1 = time.perf_counter()
        run_reference(burst, cfg, K, align_only=align_only)
        t2 = time.perf_counter()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **K.d)
        report[name] = {"gen_s": t1 - t0, "reference_s": t2 - t1, "bytes": os.path.getsize(path), "keys": len(K.d),
                        "tile_sizes": list(cfg.block_matching.tuning.tile_sizes)}
        print(name, report[name], flush=True)
        del burst
        torch.cuda.empty_cache()
    json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
