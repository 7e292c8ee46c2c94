# grey-image FFT passes: parity tests, CUDA-event times against the cuFFT route, per-kernel durations (ncu launch list)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grey" 2>&1 | tail -5
timeout 300 python tools/grey_microbench.py
ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/grey_launches.csv python tools/grey_microbench.py > /dev/null 2>&1
grep "grey_" gpurun_out/grey_launches.csv | awk -F'","' '{print $5, $(NF)}' | sed 's/(.*)//' | sort | uniq -c | sort -k2,2 -k3,3n | awk '{print}' | head -60
