set -x
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
e=l["e2e"]; p=e.get("uint16_in_uint8_out") or {}
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "lat", round(e["single_burst_latency_ms"],2), "u16", round(e["uint16_raw"]["ms_per_step"],2), "post", p.get("ms_per_step"), p.get("single_burst_latency_ms"))
PY
}
HHSR_ALIGN_AHEAD=1 run a1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=2 run a2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=3 run a3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_ALIGN_AHEAD=4 run a4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
