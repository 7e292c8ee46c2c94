set -x
for a in 1 3; do
HHSR_ALIGN_AHEAD=$a python tools/phase_breakdown.py 20x12MP_s2 2>&1 | grep "^{"
HHSR_ALIGN_AHEAD=$a python tools/host_overhead.py 20
done
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
e=l["e2e"]; p=e.get("uint16_in_uint8_out") or {}
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "lat", round(e["single_burst_latency_ms"],2), "u16", round(e["uint16_raw"]["ms_per_step"],2), "post", p.get("ms_per_step"), p.get("single_burst_latency_ms"), "ms/frame", round(l["roofline"]["ms_per_frame"],4))
PY
}
HHSR_ALIGN_AHEAD=1 run f1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=3 run f3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=1 run f1b python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=3 run f3b python bench.py --steps 10 --warmup 3 --no-cpu-baseline
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py -m gpu -x -q 2>&1 | tail -3
