set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ica or bm_l1" 2>&1 | tail -3
timeout 300 python tools/e2e_steps.py 5
timeout 300 python tools/e2e_steps.py 19
