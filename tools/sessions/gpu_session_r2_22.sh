set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_$1.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("$1 ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3), round(e["uint16_in_uint8_out"]["single_burst_latency_ms"],3))
PY
}
for B in 5 10 19; do
timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline --merge-batch $B > gpurun_out/bench_r2_pop$B.json 2> gpurun_out/bench_r2_pop$B.err; show pop$B
done
HHSR_STAGING_SLOTS=24 timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline --merge-batch 19 > gpurun_out/bench_r2_pop19s24.json 2> gpurun_out/bench_r2_pop19s24.err; show pop19s24
timeout 300 python tools/e2e_timeline.py 2>&1 | grep -v "^  start" | tail -30
