set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/stage_microbench.py --only robustness,align,grey_fft --iters 30
for v in variants/rob_mb6.so variants/rob_mb5.so; do echo $v; HHSR_LIB=$PWD/$v python tools/stage_microbench.py --only robustness --iters 30; done
ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rob_launches.csv python tools/stage_microbench.py --iters 3 --only robustness > /dev/null 2>&1
grep "robustness_kernel\|local_min5\|guide_stats" gpurun_out/rob_launches.csv | awk -F'","' '{print $5, $(NF)}' | sed 's/(.*)//' | sort | uniq -c | sort -k2,2 -k3,3n | head
