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


if __name__ == "__main__":
    launches()
    merge()
