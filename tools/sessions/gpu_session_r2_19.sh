set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/stage_microbench.py --iters 20 --only grey_fft,pyramid,align,robustness,estimate_kernels
ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/stage_launches.csv python tools/stage_microbench.py --iters 2 --only pyramid,align > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/stage_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0][:70]].append(float(r[-1]) / 1e3)
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print("%-72s n=%3d min %.1f med %.1f us" % (k, len(v), min(v), sorted(v)[len(v)//2]))
PY
