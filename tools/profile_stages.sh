#!/bin/bash
# ncu --set full capture of ONE launch of each hot kernel (driver: tools/stage_microbench.py, one 12 MP frame), raw
# metric pages exported as CSV into gpurun_out/ncu_<tag>/.  Usage: tools/profile_stages.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out/ncu_$TAG
mkdir -p $OUT
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(accumulate_kernel robustness_kernel local_min5 bm_l2_tiled32 "ica_kernel<32" estimate_kernels_kernel "gauss_downsample_kernel<2" guide_stats grey_band_mask accumulate_ref)
fi
for K in "${KERNELS[@]}"; do
  N=$(echo "$K" | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 3 -c 1 -f -o $OUT/$N \
      python tools/stage_microbench.py --iters 2 > $OUT/$N.log 2>&1
  ncu -i $OUT/$N.ncu-rep --page raw --csv > $OUT/$N.raw.csv 2>/dev/null
done
ls -la $OUT
