"""Turn the raw ncu exports under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarise_profiles.py r01
reads  gpurun_out/launches_<tag>.csv          (ncu --metrics gpu__time_duration.sum --csv launch list of bench.py)
       gpurun_out/merge_<tag>_full.raw.csv    (ncu -i merge_<tag>_full.ncu-rep --page raw --csv, one merge launch)
writes profiles/launches_<tag>_summary.md, profiles/merge_accumulate_<tag>_ncu.md, profiles/merge_traffic_bytes.json
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3      # ns -> us
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "hhsr::" in k)
    out = ["# ncu launch list — `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e` (20x12MP_s2, 1xB200), round %s" % tag[1:],
           "",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_%s.csv "
           "python bench.py --steps 1 --warmup 3 --no-cpu-baseline`" % tag,
           "(warm-up steps, the timed resident step, the burst generation are all in the list; "
           "per-launch times are cold-cache and serialised: compare SHARES).", "",
           "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:34]:
        out.append("| `%s` | %d | %.2f | %.1f | %.1f%% |" % (k[:70], n, t / 1e3, t / n, 100 * t / total))
    out += ["", "Total GPU time in list: %.1f ms; libhhsr kernels: %.1f ms (%.1f%%); the rest is cuFFT (grey image), torch "
            "fills/copies and the synthetic-burst generator." % (total / 1e3, ours / 1e3, 100 * ours / total)]
    m = [(k, v) for k, v in agg.items() if "accumulate_pow2_kernel" in k and "(bool)0>" in k.replace(" ", "")] or \
        [(k, v) for k, v in agg.items() if "accumulate_pow2_kernel" in k]
    for k, (n, t) in m:
        out.append("")
        out.append("`%s`: %d launches, avg %.1f us under ncu; share of all libhhsr time: %.1f%%." % (k[:60], n, t / n, 100 * t / ours))
    open(os.path.join(P, "launches_%s_summary.md" % tag), "w").write("\n".join(out) + "\n")


def merge():
    rows = list(csv.reader(open(os.path.join(G, "merge_%s_full.raw.csv" % tag))))
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    name = d["Kernel Name"][1]
    out = ["# ncu --set full: `%s` (merge of one 12 MP comp frame into the 48 MP accumulators), round %s" % (name[:80], tag[1:]), "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:accumulate_pow2 -s 40 -c 1 -o "
           "gpurun_out/merge_%s_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline`" % tag, "",
           "| metric | value | unit |", "|---|---:|---|"]
    for k in keys:
        if k in d:
            out.append("| %s | %s | %s |" % (k, d[k][1], d[k][0]))

    def gb(k):
        u, v = d[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    out += ["", "DRAM traffic per launch: %.3f GB (read %.3f + write %.3f) vs algorithmic 2.448 GB (48 B per HR pixel + 12 B per "
            "LR pixel): no re-reads." % ((rd + wr) / 1e9, rd / 1e9, wr / 1e9)]
    open(os.path.join(P, "merge_accumulate_%s_ncu.md" % tag), "w").write("\n".join(out) + "\n")
    json.dump({"20x12MP_s2": rd + wr, "8x12MP_s2": rd + wr,
               "source": "profiles/merge_accumulate_%s_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)" % tag},
              open(os.path.join(P, "merge_traffic_bytes.json"), "w"))


# algorithmic bytes of one launch on a 12 MP frame at scale 2: what the stage must move at least
ALGORITHMIC = {
    "merge_finish": ("merge, whole burst: 19 comp frames + reference frame + divide in one pass (round 2b)", 48e6 * 12 + 19 * (12e6 * 12 + 94 * 125 * 8) + 12e6 * 8),
    "accumulate_pow2_batch": ("merge, 4 frames per pass", 48e6 * 48 + 4 * 12e6 * 12),
    "accumulate_pow2_kernel": ("merge, 1 frame per pass", 48e6 * 48 + 12e6 * 12),
    "robustness_kernel": ("fused robustness", (7 * 48 + 36 + 48) * 1e6),
    "local_min5": ("5x5 minimum", 96e6),
    "bm_l2_tiled32": ("L2 block matching, level 1 (2852 tiles)", 2852 * (40 * 40 + 32 * 32) * 4),
    "ica32_grad": ("ICA level 0 (11750 tiles), gradients re-formed in the kernel (round 2b)", 2 * 48e6),
    "ica32_kernel": ("ICA level 0 (11750 tiles)", 4 * 48e6),
    "estimate_kernels_kernel": ("kernel estimation", 96e6),
    "gauss_downsample_stream": ("pyramid level 0 -> 1, column-streaming kernel (round 2b)", 48e6 + 12e6),
    "gauss_downsample": ("pyramid level 0 -> 1", 48e6 + 12e6),
    "grey_rows_forward": ("grey image pass 1: row pairs forward, pruned store (round 2b)", 48e6 + 3000 * 1008 * 8),
    "grey_cols": ("grey image pass 2: columns forward + band mask + inverse (round 2b)", 2 * 3000 * 1008 * 8),
    "grey_rows_inverse": ("grey image pass 3: rows inverse (round 2b)", 3000 * 1008 * 8 + 48e6),
    "guide_stats": ("guide image + local stats", 48e6 + 36e6),
    "grey_band_mask": ("band mask on the half spectrum", 48e6),
    "accumulate_ref": ("merge_ref + divide", 48e6 * 48 + 12e6 * 8),
    "regular_fft": ("cuFFT column pass (library)", 2 * 96e6),
    "vector_fft": ("cuFFT row pass (library)", 48e6 + 96e6),
    "post_blur_cols": ("unsharp mask, column pass (48 MP x 3)", 2 * 576e6),
    "post_finish": ("unsharp mask row pass + gamma + uint8", 2 * 576e6 + 144e6),
}


def stages():
    """profiles/stages_<tag>_ncu.md from gpurun_out/ncu_<tag>/*.raw.csv (tools/profile_stages.sh)."""
    d = os.path.join(G, "ncu_%s" % tag)
    if not os.path.isdir(d):
        return
    out = ["# ncu --set full, one launch per hot kernel — `tools/profile_stages.sh %s` (driver: tools/stage_microbench.py, one 12 MP "
           "frame, scale 2, 1xB200), round %s" % (tag, tag[1:]), "",
           "Command per kernel: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 3 -c 1 python "
           "tools/stage_microbench.py --iters 2`.  `DRAM bytes` = dram__bytes_read.sum + dram__bytes_write.sum of that launch; "
           "`algorithmic` = the bytes the stage must move at least (DESIGN.md section 4); `HBM frac` = algorithmic bytes / duration / "
           "6536 GB/s (MEASURED_PEAKS.json).", "",
           "| kernel | what | us | DRAM bytes (MB) | algorithmic (MB) | HBM frac | DRAM % | SM % | issue % | warps % | regs | L1 hit % | L2 hit % |",
           "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]

    def num(x):
        return float(x.replace(",", ""))
    for f in sorted(os.listdir(d)):
        if not f.endswith(".raw.csv"):
            continue
        rows = list(csv.reader(open(os.path.join(d, f))))
        if len(rows) < 3:
            continue
        m = dict(zip(rows[0], zip(rows[1], rows[2])))

        def val(k, default=float("nan")):
            if k not in m:
                return default
            u, v = m[k]
            v = num(v)
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
        key = f[:-8]
        what, alg = next(((w, a) for k, (w, a) in ALGORITHMIC.items() if key.startswith(k)), ("", float("nan")))
        us = val("gpu__time_duration.sum")
        dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        out.append("| `%s` | %s | %.1f | %.0f | %.0f | %.2f | %.0f | %.0f | %.0f | %.0f | %d | %.0f | %.0f |" % (
            m["Kernel Name"][1][:48], what, us, dram / 1e6, alg / 1e6, alg / (us * 1e-6) / 6536.4e9,
            val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), val("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            val("smsp__issue_active.avg.pct_of_peak_sustained_active"), val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            int(val("launch__registers_per_thread", 0)), val("l1tex__t_sector_hit_rate.pct"), val("lts__t_sector_hit_rate.pct")))
    open(os.path.join(P, "stages_%s_ncu.md" % tag), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    for fn in (launches, merge, stages):
        try:
            fn()
        except FileNotFoundError as e:
            print("skipped %s: %s" % (fn.__name__, e))
