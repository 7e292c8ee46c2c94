#!/bin/bash
# ncu --set full of the kernels written in the second half of round 2 (grey-image FFT passes, streaming pyramid kernel,
# ICA with in-kernel gradients) + the launch list of the bench command.  Raw CSV pages into gpurun_out/ncu_r02b/.
set -u
OUT=gpurun_out/ncu_r02b
mkdir -p $OUT
cap() {   # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $OUT/$1 \
      python tools/stage_microbench.py --iters 2 > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
}
cap grey_rows_forward grey_rows_forward_kernel 2
cap grey_cols grey_cols_kernel 2
cap grey_rows_inverse grey_rows_inverse_kernel 2
cap gauss_downsample_stream gauss_downsample_stream_kernel 3
cap ica32_grad ica32_kernel 5
ls -la $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_r02b.log 2>&1
tail -2 gpurun_out/launches_r02b.log | cut -c1-300
wc -l gpurun_out/launches_r02b.csv
