"""GPU parity AT THE BENCHMARK SHAPES (`-m gpu`): the CUDA path against outputs of the UNMODIFIED reference run on a
B200 on the very bursts bench.py times — 3000x4000 at scales 2 and 3, 6144x8192 (alignment), plus whole pipelines at
tile sizes 64 and 16 (tests/golden/make_golden_bench_gpu.py wrote the fixtures; the bursts are regenerated from the
seeded generator and proven identical through their stored float64 sums and a crop).

What is asserted
  * block matching: the flow after EVERY L2 level equals the reference's bit for bit on EVERY tile, when fed the
    reference's own previous-level flow.  The mismatch count is reported and must be 0 on the benchmark bursts; elsewhere
    a mismatch is only accepted as a proven near-tie: our float64 SSD of our offset is not larger than that of the
    reference's offset and the two differ by less than 1e-4 (the reference's float32 FFT correlation cannot resolve it);
  * ICA per level < 2e-5 px on every tile from the reference's block-matching output; whole alignment chain
    (our flows carried level to level) < 1e-4 px on every tile;
  * grey image, Hessians, robustness r, covariances, accumulators and the output image on 3x3 grids of crops
    (corners, edges, centre) within the tolerances of tests/test_gpu_parity.py, float64 sums of the full arrays to
    1e-6 relative, identical NaN counts.
The numbers measured by a run are written to gpurun_out/bench_shape_parity_report.json (a copy is tracked under
profiles/)."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, attr_cfg, load

pytestmark = pytest.mark.gpu

REPORT = {}
PIPELINE_CASES = ["bench12_s2", "bench12_s3", "ts64_pipeline", "ts16_pipeline"]
ALIGN_CASES = PIPELINE_CASES + ["bench50_align"]


def record(case, name, value):
    REPORT.setdefault(case, {})[name] = value
    out = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(REPORT, open(os.path.join(out, "bench_shape_parity_report.json"), "w"), indent=1, sort_keys=True)


def _generator_module():
    spec = importlib.util.spec_from_file_location("hhsr_crops", os.path.join(GOLDEN, "crop_grid.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


CG = _generator_module()
_CACHE = {}


def get_case(name):
    """Fixture + regenerated burst (device tensor) + config; cached one case at a time (bursts are up to 400 MB)."""
    if name in _CACHE:
        return _CACHE[name]
    _CACHE.clear()
    torch.cuda.empty_cache()
    from handheld_super_resolution.synthetic import synth_burst
    z = load(name + ".npz")
    c = json.loads(str(z["case"]))
    burst, _ = synth_burst(c["n"], c["H"], c["W"], seed=c["seed"], device="cuda", as_numpy=False)
    sums = burst.double().sum(dim=(1, 2)).cpu().numpy()
    same = np.array_equal(sums, z["burst__sum"]) and np.array_equal(burst[-1][100:132, 200:232].cpu().numpy(), z["burst__crop"])
    assert same, ("the seeded generator did not reproduce the burst the goldens of %s were made from (sums %s vs %s): "
                  "regenerate the fixtures with tests/golden/make_golden_bench_gpu.py on this software stack" % (name, sums, z["burst__sum"]))
    cfg = attr_cfg(z["cfg_json"])
    _CACHE[name] = (z, c, burst, cfg)
    return _CACHE[name]


def full_summary(t):
    """Same four numbers as the generator's summary(): finite float64 sum, NaN / zero / inf counts (on the device)."""
    fin = torch.isfinite(t)
    return np.array([torch.where(fin, t, torch.zeros((), dtype=t.dtype, device=t.device)).double().sum().item(),
                     torch.isnan(t).sum().item(), (t == 0).sum().item(), torch.isinf(t).sum().item()])


def crops_of(t, size):
    return np.stack([t[y:y + size, x:x + size].cpu().numpy() for y, x in CG.crop_origins(tuple(t.shape), size)])


def crop_diff(t, want_crops, rel_floor=None):
    """max |ours - reference| over the 3x3 crop grid (entries finite in both); NaN patterns must coincide.
    rel_floor: compare relative to max(|reference|, rel_floor) instead."""
    size = want_crops.shape[1]
    got = crops_of(t, size).astype(np.float64)
    want = want_crops.astype(np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs inside the crops"
    m = np.isfinite(got) & np.isfinite(want)
    d = np.abs(got[m] - want[m])
    if rel_floor is not None:
        d = d / np.maximum(np.abs(want[m]), rel_floor)
    return float(d.max()) if d.size else 0.0


def check_summary(case, name, t, want, sum_rtol=1e-6, zero_slack=0):
    got = full_summary(t)
    record(case, name + "__sum", {"ours": got.tolist(), "reference": want.tolist()})
    assert abs(got[0] - want[0]) <= sum_rtol * max(abs(want[0]), 1.0), (name, got, want)
    assert got[1] == want[1], "%s: NaN count %d vs %d" % (name, got[1], want[1])
    assert got[3] == want[3], "%s: inf count %d vs %d" % (name, got[3], want[3])
    assert abs(got[2] - want[2]) <= zero_slack, "%s: zero count %d vs %d" % (name, got[2], want[2])


def ssd_margin(ref_lvl, mov_lvl, ts, ty, tx, flow_in, ours, theirs):
    """float64 SSD (block_matching.py:20-76 semantics, clamped reads) of the two candidate offsets of one tile."""
    def ssd(off):
        fx, fy = int(round(float(flow_in[0]))) + int(off[0]), int(round(float(flow_in[1]))) + int(off[1])
        ys = torch.clamp(torch.arange(ty * ts, (ty + 1) * ts, device="cuda") + fy, 0, mov_lvl.shape[0] - 1)
        xs = torch.clamp(torch.arange(tx * ts, (tx + 1) * ts, device="cuda") + fx, 0, mov_lvl.shape[1] - 1)
        m = mov_lvl[ys][:, xs].double()
        r = ref_lvl[ty * ts:(ty + 1) * ts, tx * ts:(tx + 1) * ts].double()
        return float((m * m - 2 * r * m).sum())
    return ssd(ours) - ssd(theirs)


@pytest.mark.parametrize("name", ALIGN_CASES)
def test_alignment_every_tile(name):
    from handheld_super_resolution import alignment as AL
    from handheld_super_resolution.utils_image import compute_grey_images
    z, c, burst, cfg = get_case(name)
    bm = cfg.block_matching.tuning
    L = len(bm.factors)
    ref_grey = compute_grey_images(burst[0], "FFT")
    d = crop_diff(ref_grey, z["grey_0__crops"])
    record(name, "grey_0_crops", d)
    assert d < 2e-6
    check_summary(name, "grey_0", ref_grey, z["grey_0__sum"])
    ref = AL.init_alignment(ref_grey, cfg)
    for i in range(L - 1):                # coarse levels: full Hessians; finest: crops + sum
        want = z["ref_hessian_c%d" % i]
        got = ref[5][i].cpu().numpy()
        rel = float((np.abs(got - want) / np.maximum(np.abs(want), 1.0)).max())
        record(name, "hessian_rel_c%d" % i, rel)
        assert rel < 2e-5                 # parallel vs sequential float32 sum of ts^2 products
    hfin = ref[5][L - 1]
    assert crop_diff(hfin.reshape(*hfin.shape[:2], 4), z["ref_hessian_c%d__crops" % (L - 1)].reshape(9, 16, 16, 4), rel_floor=1.0) < 2e-5
    total_tiles = mismatches = 0
    worst_ica = worst_e2e = 0.0
    e2e_bad_tiles = e2e_tiles = 0
    margins = []
    for f in range(1, c["n"]):
        grey = compute_grey_images(burst[f], "FFT")
        if f == 1:
            assert crop_diff(grey, z["grey_1__crops"]) < 2e-6
        mpyr = AL.build_gaussian_pyramid(grey, bm.factors)
        prev = None
        for i in range(L):
            l = L - 1 - i
            npatchs = tuple(ref[5][i].shape[:2])
            if prev is None:
                flow = torch.zeros((*npatchs, 2), device="cuda")
            else:
                flow = AL.upscale_lvl(prev, npatchs, l, cfg)
            flow_in = flow.clone()
            if bm.metrics[l] == "L2":
                AL.align_lvl_block_matching_L2(ref[1][i], ref[2][i], mpyr[i], flow, l, cfg)
                want = torch.from_numpy(z["flow_f%d_l%d_bm" % (f, l)]).cuda()
                bad = (flow != want).any(dim=-1)
                total_tiles += bad.numel()
                nbad = int(bad.sum().item())
                if nbad:
                    mismatches += nbad
                    ts = bm.tile_sizes[l]
                    for ty, tx in bad.nonzero().cpu().tolist()[:64]:
                        margin = ssd_margin(ref[1][i], mpyr[i], ts, ty, tx, flow_in[ty, tx].cpu().numpy(),
                                            (flow[ty, tx] - flow_in[ty, tx]).cpu().numpy(), (want[ty, tx] - flow_in[ty, tx]).cpu().numpy())
                        record(name, "bm_mismatch_f%d_l%d_%d_%d" % (f, l, ty, tx),
                               {"ours": flow[ty, tx].tolist(), "reference": want[ty, tx].tolist(), "ssd_ours_minus_reference": margin})
                        margins.append(margin)
                flow = want.clone()
            else:
                assert bool(z["flow_f%d_l%d_bm_is_rint" % (f, l)]), "the reference's L1 level was not rint(flow) (SURVEY Q1)"
                AL.align_lvl_block_matching_L1(ref[0][i], mpyr[i], flow, l, cfg)
                assert torch.equal(flow, torch.round(flow_in))
            AL.align_lvl_ica(ref[0][i], ref[3][i], ref[4][i], ref[5][i], mpyr[i], flow, l, cfg)
            want = torch.from_numpy(z["flow_f%d_l%d_ica" % (f, l)]).cuda()
            worst_ica = max(worst_ica, float((flow - want).abs().max().item()))
            prev = want
        e2e = AL.align(*ref, grey, cfg)
        de = (e2e - prev).abs().amax(dim=-1)
        worst_e2e = max(worst_e2e, float(de.max().item()))
        e2e_bad_tiles += int((de >= 1e-4).sum().item())
        e2e_tiles += de.numel()
    record(name, "bm_tiles_compared", total_tiles)
    record(name, "bm_offset_mismatches", mismatches)
    record(name, "ica_max_abs_px", worst_ica)
    record(name, "flow_end_to_end_max_abs_px", worst_e2e)
    record(name, "flow_end_to_end_tiles_over_1e-4px", [e2e_bad_tiles, e2e_tiles])
    assert worst_ica < 2e-5
    if name.startswith("bench"):
        assert mismatches == 0, "%d of %d tiles: block-matching offset differs from the reference" % (mismatches, total_tiles)
    else:   # near-ties only, and few of them
        assert mismatches <= 64 and mismatches <= 1e-3 * total_tiles and len(margins) == mismatches
        assert all(-1e-4 < m <= 0.0 for m in margins), margins
    if mismatches == 0:
        assert worst_e2e < 1e-4
    else:   # a near-tie decided the other way moves that tile (and what is upsampled from it) by a pixel
        assert e2e_bad_tiles <= 16 * mismatches


@pytest.mark.parametrize("name", PIPELINE_CASES)
@pytest.mark.parametrize("batch", [1, 3])
def test_pipeline_against_reference(name, batch):
    """main() on the benchmark burst, every stage output hooked and compared with the reference's."""
    from handheld_super_resolution import super_resolution as SR
    if batch != 1 and name != "bench12_s2":
        pytest.skip("the batched merge is compared on the main benchmark case")
    z, c, burst, cfg = get_case(name)
    tag = name if batch == 1 else name + "_batch%d" % batch
    # the Ts-16 / Ts-64 bursts may hold a block-matching near-tie decided the other way (test_alignment_every_tile): one tile
    # of such a frame then lands a pixel away, which moves the whole-array sums by more than rounding
    sum_rtol = 1e-6 if name.startswith("bench") else 2e-5
    n_comp = c["n"] - 1
    seen = {"rob": 0, "kern": 0, "frames_merged": 0}
    worst = {"r": 0.0, "covs": 0.0}
    saved = {k: getattr(SR, k) for k in ("compute_robustness", "estimate_kernels", "merge", "merge_batch", "merge_ref")}

    def rob(*a, **k):
        r = saved["compute_robustness"](*a, **k)
        seen["rob"] += 1
        f = seen["rob"]
        d = crop_diff(r, z["r_f%d__crops" % f])
        worst["r"] = max(worst["r"], d)
        # r in [0, 1]; thresholded pixels (exact zeros) may flip where the reference sits on the threshold
        check_summary(tag, "r_f%d" % f, r, z["r_f%d__sum" % f], sum_rtol=sum_rtol, zero_slack=max(8, int(2e-6 * r.numel())))
        return r

    def kern(img, config, **kw):
        covs = saved["estimate_kernels"](img, config, **kw)
        seen["kern"] += 1
        k = seen["kern"]
        d = crop_diff(covs.reshape(*covs.shape[:2], 4), z["covs_%d__crops" % k].reshape(9, *z["covs_%d__crops" % k].shape[1:3], 4))
        worst["covs"] = max(worst["covs"], d)
        check_summary(tag, "covs_%d" % k, covs, z["covs_%d__sum" % k])
        return covs

    def after_merge(num, den, n_frames):
        seen["frames_merged"] += n_frames
        for key, when in (("first", 1), ("comp", n_comp)):
            if seen["frames_merged"] == when and ("num_%s__crops" % key) in z:
                dn = crop_diff(num, z["num_%s__crops" % key], rel_floor=1.0)
                dd = crop_diff(den, z["den_%s__crops" % key], rel_floor=1.0)
                record(tag, "num_den_%s_crops_rel" % key, [dn, dd])
                assert dn < 5e-5 and dd < 5e-5      # float32 weights on flows that differ by a few 1e-6 px
                check_summary(tag, "num_" + key, num, z["num_%s__sum" % key], sum_rtol=sum_rtol, zero_slack=int(1e-5 * num.numel()))
                check_summary(tag, "den_" + key, den, z["den_%s__sum" % key], sum_rtol=sum_rtol, zero_slack=int(1e-5 * num.numel()))

    def merge(comp, al, covs, r, num, den, cfa, config, init=False):
        saved["merge"](comp, al, covs, r, num, den, cfa, config, init=init)
        after_merge(num, den, 1)

    def merge_batch(comps, als, covs, rs, num, den, cfa, config, **kw):
        saved["merge_batch"](comps, als, covs, rs, num, den, cfa, config, **kw)
        if kw.get("finish") is None:            # a fused finish leaves the image, not the accumulators, in num
            after_merge(num, den, len(comps))
        else:
            seen["frames_merged"] += len(comps)

    # batch == 1: the reference's structure (one merge per frame, separate merge_ref + divide), every stage hooked;
    # batch == 3: what main() does by default (batched merge, the last batch fused with merge_ref + divide): the stage
    # hooks would see the reference frame's covariances first, so only the merges are hooked
    fused = batch != 1
    hooks = (("merge", merge), ("merge_batch", merge_batch)) if fused else \
        (("compute_robustness", rob), ("estimate_kernels", kern), ("merge", merge), ("merge_batch", merge_batch))
    for k, fn in hooks:
        setattr(SR, k, fn)
    saved_fuse, SR.FUSE_FINISH = SR.FUSE_FINISH, fused
    try:
        out, dbg = SR.main(burst[0], burst[1:], cfg, merge_batch_size=batch)
        torch.cuda.synchronize()
    finally:
        SR.FUSE_FINISH = saved_fuse
        for k, fn in saved.items():
            setattr(SR, k, fn)
    assert seen["frames_merged"] == n_comp and (fused or (seen["rob"] == n_comp and seen["kern"] == c["n"]))
    if not fused:
        record(tag, "r_crops_max_abs", worst["r"])
        record(tag, "covs_crops_max_abs", worst["covs"])
        assert worst["r"] < 2e-5            # R, r in [0, 1]
        assert worst["covs"] < 1e-6
    d = crop_diff(out, z["out__crops"])
    record(tag, "out_crops_max_abs", d)
    assert d < 1e-4
    check_summary(tag, "out", out, z["out__sum"], sum_rtol=sum_rtol, zero_slack=8)
    acc = dbg["accumulated robustness"]
    assert crop_diff(acc, z["acc_rob__crops"]) < 2e-5 * n_comp
    check_summary(tag, "acc_rob", acc, z["acc_rob__sum"], sum_rtol=sum_rtol, zero_slack=max(8, int(2e-6 * acc.numel())))
