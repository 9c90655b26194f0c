#!/bin/bash
# kNN-only profile: launch list of scripts/bench_knn.py and one --set full capture of the tensor-core filtered kernel.
R=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn|sqnorm' -c 120 --csv --log-file gpurun_out/${R}_knn_launches.csv \
    python scripts/bench_knn.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 30 -c 1 -o gpurun_out/${R}_prof_knn_tc \
    python scripts/bench_knn.py > /dev/null 2>&1
ncu -i gpurun_out/${R}_prof_knn_tc.ncu-rep --page raw --csv > gpurun_out/${R}_raw_knn_tc.csv 2>/dev/null
ncu -i gpurun_out/${R}_prof_knn_tc.ncu-rep --page source --csv > gpurun_out/${R}_src_knn_tc.csv 2>/dev/null
ls -la gpurun_out | tail -8
