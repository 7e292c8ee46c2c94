#!/bin/bash
# ncu --set full capture of ONE launch of each hot kernel (driver: tools/stage_microbench.py, one 12 MP frame), raw
# metric pages exported as CSV into gpurun_out/ncu_<tag>/.  Usage: tools/profile_stages.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out/ncu_$TAG
mkdir -p $OUT
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(accumulate_pow2_batch accumulate_pow2_kernel robustness_kernel local_min5 bm_l2_tiled32 ica32_kernel estimate_kernels_kernel gauss_downsample guide_stats grey_band_mask accumulate_ref regular_fft vector_fft post_blur_cols post_finish)
fi
for K in "${KERNELS[@]}"; do
  N=$(echo "$K" | tr -c 'a-zA-Z0-9_\n' '_')
  SKIP=3; [ "$K" = ica32_kernel ] && SKIP=5      # the level-0 launch of the second alignment chain (11750 tiles)
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c 1 -f -o $OUT/$N \
      python tools/stage_microbench.py --iters 2 > $OUT/$N.log 2>&1
  ncu -i $OUT/$N.ncu-rep --page raw --csv > $OUT/$N.raw.csv 2>/dev/null
done
ls -la $OUT
