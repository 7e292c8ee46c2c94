for i in 1 2 3 4 5 6; do
HHSR_BENCH_TRACE=1 timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_ff$i.json 2> gpurun_out/bench_r2_ff$i.err
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_ff$i.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("ff$i ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3))
PY
done
python tools/host_overhead.py 2>&1 | tail -2
