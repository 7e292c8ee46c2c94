set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "roof", round(l["roofline"]["frac"],3), "ms/frame", round(l["roofline"]["ms_per_frame"],3), "launches", l["gpu_launches"], "pw", l["clocks"]["power_w_max"])
PY
}
run auto1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run auto2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run auto3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
HHSR_STAGING_SLOTS=20 run auto_slots20 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
HHSR_STAGING_SLOTS=6 run auto_slots6 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run b4 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 4
run b7 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 7
run b10 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 10
