set -x
for B in 24 10 7; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --merge-batch $B > gpurun_out/bench_r2_mb$B.json 2> gpurun_out/bench_r2_mb$B.err; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_mb$B.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("B=$B ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3), round(e["uint16_in_uint8_out"]["single_burst_latency_ms"],3))
PY
done
timeout 300 python tools/e2e_timeline.py 2>&1 | tail -40
