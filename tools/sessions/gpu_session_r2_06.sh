set -x
show() { python - <<PY
import json
txt=[x for x in open("gpurun_out/$1").read().splitlines() if x.startswith("{")]
l=json.loads(txt[-1])
print("$1", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single",{}).get("max_abs_diff"), l.get("allocator"))
for e in l.get("extra_workloads", []): print("   extra", e)
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/phase_breakdown.py 20x12MP_s2 rows 2>&1 | grep "^{"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_n8_rows.json 2> gpurun_out/bench_r2_n8_rows.err || tail -30 gpurun_out/bench_r2_n8_rows.err
show bench_r2_n8_rows.json
