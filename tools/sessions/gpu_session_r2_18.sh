set -x
mkdir -p gpurun_out/ncu_grey
for K in grey_rows_forward_kernel grey_cols_kernel grey_rows_inverse_kernel; do
  ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 2 -c 1 -f -o gpurun_out/ncu_grey/$K python tools/grey_microbench.py > gpurun_out/ncu_grey/$K.log 2>&1
  ncu -i gpurun_out/ncu_grey/$K.ncu-rep --page raw --csv > gpurun_out/ncu_grey/$K.raw.csv 2>/dev/null
done
ls -la gpurun_out/ncu_grey
