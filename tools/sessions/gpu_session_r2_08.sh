set -x
timeout 1700 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15
cat gpurun_out/stage_swap_report.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/phase_breakdown.py 20x12MP_s2 rows 2>&1 | grep "^{"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2_rows.json 2> gpurun_out/bench_r2_n2_rows.err || tail -20 gpurun_out/bench_r2_n2_rows.err
python - <<PY
import json
txt=[x for x in open("gpurun_out/bench_r2_n2_rows.json").read().splitlines() if x.startswith("{")]
l=json.loads(txt[-1])
print("rows n2", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single"))
PY
