set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grey" 2>&1 | tail -5
timeout 300 python tools/grey_microbench.py
ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/grey_launches.csv python tools/grey_microbench.py > /dev/null 2>&1
grep -i "grey_\|fft" gpurun_out/grey_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -40
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_g1.json 2> gpurun_out/bench_r2_g1.err; python - <<'PY'
import json
l=[x for x in open("gpurun_out/bench_r2_g1.json").read().splitlines() if x.startswith("{")]
d=json.loads(l[-1]); print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "lat", d["e2e"]["single_burst_latency_ms"])
PY
