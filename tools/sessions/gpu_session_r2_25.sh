set -x
show() { python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_$1.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("$1 ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3), round(e["uint16_in_uint8_out"]["single_burst_latency_ms"],3), d["clocks"])
PY
}
for i in 1 2 3; do
timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_nvml$i.json 2> gpurun_out/bench_r2_nvml$i.err; show nvml$i
done
for i in 1 2; do
HHSR_BENCH_SAMPLER=smi timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_smi$i.json 2> gpurun_out/bench_r2_smi$i.err; show smi$i
done
HHSR_BENCH_SAMPLER=off timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_off1.json 2> gpurun_out/bench_r2_off1.err; show off1
