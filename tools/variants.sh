python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -30
python tools/stage_microbench.py
for v in variants/*.so; do echo $v; HHSR_LIB=$PWD/$v python tools/stage_microbench.py --only merge --iters 20; done
