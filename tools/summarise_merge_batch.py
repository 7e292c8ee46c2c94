"""profiles/merge_batch_<tag>_ncu.md + profiles/merge_traffic_bytes.json from the raw ncu exports:
    gpurun_out/merge_<tag>_batch.raw.csv           whole-burst launch of accumulate_pow2_batch_kernel (bench.py under ncu)
    gpurun_out/ncu_<tag>/accumulate_pow2_batch.*   4-frame launch (tools/profile_stages.sh)
    gpurun_out/ncu_<tag>/accumulate_pow2_kernel.*  single-frame kernel
    python tools/summarise_merge_batch.py r02"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def load(path):
    rows = list(csv.reader(open(path)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def gb(d, k):
    u, v = d[k]
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


def main():
    dw = load(os.path.join(G, "merge_%s_batch.raw.csv" % tag))
    d4 = load(os.path.join(G, "ncu_%s" % tag, "accumulate_pow2_batch.raw.csv"))
    d1 = load(os.path.join(G, "ncu_%s" % tag, "accumulate_pow2_kernel.raw.csv"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    name = dw["Kernel Name"][1]
    fused = "(bool)1>" in name.replace(" ", "")[-12:] or name.rstrip(")").endswith("1>(")
    out = ["# ncu --set full: the merge kernels, round %s (20x12MP_s2, 1xB200)" % tag[1:], "",
           "Three captures.  (1) `%s` merging the WHOLE burst (19 comp frames, initialising%s) in one pass - what `main()` launches "
           "on a resident burst: `ncu --set full --clock-control none --import-source on -k regex:accumulate_pow2_batch -s 3 -c 1 "
           "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e`.  (2) The same kernel family on 4 frames, accumulating "
           "(`tools/profile_stages.sh %s`).  (3) The single-frame `accumulate_pow2_kernel` (round 1's kernel, the reference's launch "
           "granularity)." % (name[:60], ", fused with merge_ref + divide" if fused else "", tag), "",
           "| metric | whole burst | 4 frames | single frame |", "|---|---:|---:|---:|"]

    def cell(d, k):
        return ("%s %s" % (d[k][1], d[k][0])) if k in d else ""
    for k in keys:
        out.append("| %s | %s | %s | %s |" % (k, cell(dw, k), cell(d4, k), cell(d1, k)))
    tw, t4, t1 = (gb(d, "dram__bytes_read.sum") + gb(d, "dram__bytes_write.sum") for d in (dw, d4, d1))
    tiles = 94 * 125 * 8
    aw = (48e6 * 12 + 12e6 * 8 if fused else 48e6 * 24) + 19 * (12e6 * 12 + tiles)
    a4, a1 = 48e6 * 48 + 4 * (12e6 * 12 + tiles), 48e6 * 48 + 12e6 * 12 + tiles
    out += ["", "DRAM traffic per launch against the algorithmic bytes `B_batch(K) = HR*24*(1+[not init]) + K*(LR*12 + tiles*8)` "
            "(a finishing launch writes HR*12 - the image - instead of HR*24 and reads LR*8 more):", "",
            "| launch | DRAM (GB) | algorithmic (GB) | ratio | DRAM per frame (GB) |", "|---|---:|---:|---:|---:|",
            "| 19 frames, initialising%s | %.3f | %.3f | %.2f | %.3f |" % (" + finish" if fused else "", tw / 1e9, aw / 1e9, tw / aw, tw / 19e9),
            "| 4 frames, accumulating | %.3f | %.3f | %.2f | %.3f |" % (t4 / 1e9, a4 / 1e9, t4 / a4, t4 / 4e9),
            "| 1 frame, accumulating (single-frame kernel) | %.3f | %.3f | %.2f | %.3f |" % (t1 / 1e9, a1 / 1e9, t1 / a1, t1 / 1e9),
            "", "No wasted re-reads (traffic <= algorithmic bytes; part of the LR planes of a 4-frame launch is still in the 126 MB L2). "
            "The single-frame kernel is HBM-bound (DRAM ~62 %, 0.80 of the measured copy peak); the batched kernel moves 6-19x fewer "
            "bytes per frame and is bound by instruction issue (~75 %)."]
    rep = os.path.join(G, "ncu_%s" % tag, "accumulate_pow2_batch.ncu-rep")
    if os.path.exists(rep):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        hdr = next((r for r in rows if "Instructions Executed" in r), None)
        if hdr:
            ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
            data = [(r[ia].strip(), int(r[ie])) for r in rows if len(r) > ie and r[ie].isdigit()]
            tot = sum(e for _, e in data)
            by = collections.Counter()
            for s_, e in data:
                by[(s_.split()[1] if s_.startswith("@") else s_.split()[0]).split(".")[0]] += e
            out += ["", "Executed instruction mix of the 4-frame launch (`ncu -i ... --page source --csv --print-source sass`; "
                    "%.0f warp instructions per thread and frame):" % (tot / 375000 / 4), "", "| opcode | share | per thread and frame |",
                    "|---|---:|---:|"]
            for op, e in by.most_common(14):
                out.append("| %s | %.1f %% | %.0f |" % (op, 100 * e / tot, e / 375000 / 4))
            out += ["", "FSEL is the RGGB channel resolve (per-parity partial sums -> R, G, B), FMNMX the per-tap clamp `max(0, z)` of "
                    "merge.py:424, MUFU the 9 `ex2.approx` + the reciprocal of the determinant per pixel."]
    open(os.path.join(P, "merge_batch_%s_ncu.md" % tag), "w").write("\n".join(out) + "\n")
    json.dump({"20x12MP_s2_batch19": tw, "20x12MP_s2_batch4": t4, "20x12MP_s2_batch1": t1,
               "source": "profiles/merge_batch_%s_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum of one launch each)" % tag},
              open(os.path.join(P, "merge_traffic_bytes.json"), "w"))


if __name__ == "__main__":
    main()
