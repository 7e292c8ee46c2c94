set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm_l2" -c 12 --csv --log-file gpurun_out/bm_times.csv python tools/stage_microbench.py --iters 2 --only align > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/bm_times.csv")) if len(r)>10]
h=rows[0]; ik,iv,ig=h.index("Kernel Name"),h.index("Metric Value"),h.index("Grid Size")
for r in rows[1:8]: print(r[ik][:40], r[ig], r[iv])
PY
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py -m gpu -x -q 2>&1 | tail -5
