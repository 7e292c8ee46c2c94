set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
cat gpurun_out/stage_swap_report.txt
python tools/host_overhead.py 20
python tools/host_overhead.py 4
python tools/stage_microbench.py --iters 10
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_r02.log 2>&1
bash tools/profile_stages.sh r02 > gpurun_out/profile_stages_r02.log 2>&1
ls gpurun_out/ncu_r02 | head -40
