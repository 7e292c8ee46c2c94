set -x
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
python -c "
import json; l=json.load(open('gpurun_out/multi_gpu_check_n8.json')); print(len(l)); [print(x['case'],x['mode'],x['max_abs_diff'],x['ok']) for x in l]"
show() { python - <<PY
import json
txt=[x for x in open("gpurun_out/$1").read().splitlines() if x.startswith("{")]
l=json.loads(txt[-1])
print("$1", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single",{}).get("max_abs_diff"), l.get("allocator"))
for e in l.get("extra_workloads", []): print("   extra", e["workload"], round(e["ms_per_step"],2), "ms", round(e["value"]), "MPix/s")
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_n8_rows.json 2> gpurun_out/bench_r2_n8_rows.err || tail -30 gpurun_out/bench_r2_n8_rows.err
show bench_r2_n8_rows.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r2_n4_rows.json 2> gpurun_out/bench_r2_n4_rows.err || tail -30 gpurun_out/bench_r2_n4_rows.err
show bench_r2_n4_rows.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 --reduce p2p --extra-workloads none > gpurun_out/bench_r2_n8_p2p.json 2> gpurun_out/bench_r2_n8_p2p.err || tail -30 gpurun_out/bench_r2_n8_p2p.err
show bench_r2_n8_p2p.json
