set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm_l2|ica32|gauss_down" -c 40 --csv --log-file gpurun_out/bm_times.csv python tools/stage_microbench.py --iters 2 --only align > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/bm_times.csv")) if len(r)>10]
h=rows[0]; ik,iv,ig=h.index("Kernel Name"),h.index("Metric Value"),h.index("Grid Size")
for r in rows[1:14]: print(r[ik][:40], r[ig], r[iv])
PY
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py -m gpu -x -q 2>&1 | tail -5
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
e=l["e2e"]; p=e.get("uint16_in_uint8_out") or {}
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "lat", round(e["single_burst_latency_ms"],2), "u16", round(e["uint16_raw"]["ms_per_step"],2), "post", p.get("ms_per_step"), p.get("single_burst_latency_ms"), "ms/frame", round(l["roofline"]["ms_per_frame"],3), "roof", round(l["roofline"]["frac"],3))
PY
}
run d1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
