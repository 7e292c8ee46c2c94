set -x
show() { python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_$1.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("$1 ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3), round(e["uint16_in_uint8_out"]["single_burst_latency_ms"],3))
PY
}
for B in 5 10; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --merge-batch $B > gpurun_out/bench_r2_cur$B.json 2> gpurun_out/bench_r2_cur$B.err; show cur$B
done
HHSR_STAGING_SLOTS=24 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --merge-batch 19 > gpurun_out/bench_r2_cur19.json 2> gpurun_out/bench_r2_cur19.err; show cur19
HHSR_STAGING_SLOTS=24 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --merge-batch 5 > gpurun_out/bench_r2_cur5s24.json 2> gpurun_out/bench_r2_cur5s24.err; show cur5s24
timeout 300 python tools/e2e_timeline.py 2>&1 | tail -45
