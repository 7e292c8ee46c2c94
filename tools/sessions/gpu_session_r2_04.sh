set -x
nvidia-smi -L
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
cat gpurun_out/multi_gpu_check_n2.json | head -60
for mode in rows p2p reduce_scatter; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --reduce $mode > gpurun_out/bench_r2_n2_$mode.json 2> gpurun_out/bench_r2_n2_$mode.err || tail -20 gpurun_out/bench_r2_n2_$mode.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_n2_$mode.json"))
print("$mode", "ms", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "lat", round(l["e2e"]["single_burst_latency_ms"],2), "u16", round(l["e2e"]["uint16_raw"]["ms_per_step"],2), "parity", l.get("parity_vs_single"), l.get("allocator"))
PY
done
