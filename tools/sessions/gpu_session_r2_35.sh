set -x
show() { python - <<PY
import json
l=[x for x in open("gpurun_out/bench_r2_$1.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]; print("$1 ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "lat", round(e["single_burst_latency_ms"],3), "u16", round(e["uint16_raw"]["ms_per_step"],3), "u16->u8", round(e["uint16_in_uint8_out"]["ms_per_step"],3))
PY
}
for A in 2 3; do
HHSR_ALIGN_AHEAD=$A timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_ah$A.json 2> gpurun_out/bench_r2_ah$A.err; show ah$A
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload 8x12MP_s2 > gpurun_out/bench_r2_w8.json 2> gpurun_out/bench_r2_w8.err; show w8
mkdir -p gpurun_out/ncu_r02b
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:accumulate_pow2_batch_kernel" -s 3 -c 1 -f -o gpurun_out/ncu_r02b/merge_finish python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r02b/merge_finish.log 2>&1
ncu -i gpurun_out/ncu_r02b/merge_finish.ncu-rep --page raw --csv > gpurun_out/ncu_r02b/merge_finish.raw.csv 2>/dev/null
ls -la gpurun_out/ncu_r02b/merge_finish*
