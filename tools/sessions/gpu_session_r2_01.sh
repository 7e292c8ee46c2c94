set -x
python tests/golden/make_golden_bench_gpu.py > gpurun_out/golden_bench.log 2>&1 || tail -20 gpurun_out/golden_bench.log
for f in bench12_s2 bench12_s3 bench50_align ts64_pipeline ts16_pipeline; do cp gpurun_out/golden_bench/$f.npz tests/golden/; done
ls -la tests/golden/*.npz
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for b in 1 3 5 7 10 19; do python bench.py --steps 5 --warmup 3 --merge-batch $b --no-cpu-baseline > gpurun_out/bench_r2_batch$b.json 2> gpurun_out/bench_r2_batch$b.err || tail -5 gpurun_out/bench_r2_batch$b.err; python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_batch$b.json"))
print("batch", $b, "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "roof", round(l["roofline"]["frac"],3), "ms/frame", round(l["roofline"]["ms_per_frame"],3), "pf", l["roofline"]["per_frame_kernel"] and round(l["roofline"]["per_frame_kernel"]["frac"],3))
PY
done
