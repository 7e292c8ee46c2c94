set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/pcie_probe.py
run() { tag=$1; shift; "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -5 gpurun_out/bench_r2_$tag.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_$tag.json"))
print("$tag", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "roof", round(l["roofline"]["frac"],3), "ms/frame", round(l["roofline"]["ms_per_frame"],3), "launches", l["gpu_launches"])
PY
}
run b24 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
HHSR_STAGING_SLOTS=12 run b24_slots12 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
HHSR_STAGING_SLOTS=24 run b24_slots24 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run b1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 1
run b1_again python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 1
HHSR_STAGING_SLOTS=12 run b5_slots12 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --merge-batch 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_pow2_batch -s 3 -c 1 -f -o gpurun_out/merge_r02_batch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_merge_r02.log 2>&1
ncu -i gpurun_out/merge_r02_batch.ncu-rep --page raw --csv > gpurun_out/merge_r02_batch.raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
