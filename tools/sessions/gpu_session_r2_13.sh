set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
cp gpurun_out/bench_shape_parity_report.json gpurun_out/bench_shape_parity_r02.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_pow2_batch -s 3 -c 1 -f -o gpurun_out/merge_r02_batch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_merge_r02.log 2>&1
ncu -i gpurun_out/merge_r02_batch.ncu-rep --page raw --csv > gpurun_out/merge_r02_batch.raw.csv 2>/dev/null
bash tools/profile_stages.sh r02 > gpurun_out/profile_stages_r02.log 2>&1
rm -f gpurun_out/ncu_r02/post_*.ncu-rep gpurun_out/ncu_r02/regular_fft.ncu-rep gpurun_out/ncu_r02/vector_fft.ncu-rep gpurun_out/ncu_r02/accumulate_ref.ncu-rep gpurun_out/ncu_r02/local_min5.ncu-rep gpurun_out/ncu_r02/grey_band_mask.ncu-rep
python tools/stage_microbench.py --iters 10
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err || tail -5 gpurun_out/bench_r02_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err || tail -5 gpurun_out/bench_r02_reference.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_r02_n1.json"))
e=l["e2e"]; p=e.get("uint16_in_uint8_out") or {}
print("n1", "ms", round(l["ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "lat", round(e["single_burst_latency_ms"],2), "u16", round(e["uint16_raw"]["ms_per_step"],2), "post", p.get("ms_per_step"), p.get("single_burst_latency_ms"), "ms/frame", round(l["roofline"]["ms_per_frame"],3), "roof", round(l["roofline"]["frac"],3), l["roofline"]["per_frame_kernel"]["frac"], l["cpu_baseline"]["value"])
r=json.loads([x for x in open("gpurun_out/bench_r02_reference.json").read().splitlines() if x.startswith("{")][-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"], r["reference_numba_gpu"])
PY
