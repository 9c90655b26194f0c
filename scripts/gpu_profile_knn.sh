#!/bin/bash
# kNN-only profile: launch list of scripts/bench_knn.py and --set full captures of the tensor-core filtered pipeline.
R=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn|sqnorm' -c 120 --csv --log-file gpurun_out/${R}_knn_launches.csv \
    python scripts/bench_knn.py > /dev/null 2>&1
for K in knn_tc_kernel knn_finish_kernel knn_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -o gpurun_out/${R}_prof_$K \
      python scripts/bench_knn.py > /dev/null 2>&1
  ncu -i gpurun_out/${R}_prof_$K.ncu-rep --page raw --csv > gpurun_out/${R}_raw_$K.csv 2>/dev/null
  ncu -i gpurun_out/${R}_prof_$K.ncu-rep --page source --csv > gpurun_out/${R}_src_$K.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
