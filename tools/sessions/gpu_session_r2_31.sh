set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err || tail -30 gpurun_out/bench_r02_n2.err
python - <<PY
import json
txt=[x for x in open("gpurun_out/bench_r02_n2.json").read().splitlines() if x.startswith("{")]
l=json.loads(txt[-1])
print("n2 ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single",{}).get("max_abs_diff"))
PY
