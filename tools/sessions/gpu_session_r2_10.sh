set -x
timeout 1700 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15
cat gpurun_out/stage_swap_report.txt
python tools/stage_microbench.py --iters 10 --only align,pyramid,grey_fft
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
e=l["e2e"]; p=e.get("uint16_in_uint8_out") or {}
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "lat", round(e["single_burst_latency_ms"],2), "u16", round(e["uint16_raw"]["ms_per_step"],2), "post", p.get("ms_per_step"), p.get("single_burst_latency_ms"), "ms/frame", round(l["roofline"]["ms_per_frame"],3))
PY
}
run c1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run c2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run w8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload 8x12MP_s2
