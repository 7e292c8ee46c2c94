set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
show() { python - <<PY
import json
txt=[x for x in open("gpurun_out/$1").read().splitlines() if x.startswith("{")]
l=json.loads(txt[-1])
print("$1", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single",{}).get("max_abs_diff"))
for e in l.get("extra_workloads", []): print("   extra", e.get("workload"), e.get("ms_per_step"), e.get("error"))
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/phase_breakdown.py 20x12MP_s2 rows 2>&1 | grep "^{" > gpurun_out/phases_r02_n8_rows.json; cat gpurun_out/phases_r02_n8_rows.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err || tail -30 gpurun_out/bench_r02_n8.err
show bench_r02_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --reduce p2p --extra-workloads none > gpurun_out/bench_r02_n8_p2p.json 2> gpurun_out/bench_r02_n8_p2p.err || tail -30 gpurun_out/bench_r02_n8_p2p.err
show bench_r02_n8_p2p.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err || tail -30 gpurun_out/bench_r02_n4.err
show bench_r02_n4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err || tail -30 gpurun_out/bench_r02_n2.err
show bench_r02_n2.json
