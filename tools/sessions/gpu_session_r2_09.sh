set -x
timeout 1700 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15
cat gpurun_out/stage_swap_report.txt
python tools/pcie_probe.py
