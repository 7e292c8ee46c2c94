set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 600 gpurun_out/bench_r02_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 1500 gpurun_out/bench_r02_reference.json
