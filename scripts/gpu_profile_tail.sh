#!/bin/bash
# Refresh of the round profile after a change to one kernel: the ncu launch list of the bench step and a --set full capture
# of the kernels named in KERNELS (default: linear_kernel).  ncu serialises the launches, so the programmatic-launch
# attribute is switched off for these runs (GFS3D_PDL=0).
R=${1:-r2}
mkdir -p gpurun_out
export GFS3D_PDL=0
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-train --skip-kmeans > /dev/null 2>&1
for K in ${KERNELS:-linear_kernel}; do
  timeout 90 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/${R}_prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-train --skip-kmeans > /dev/null 2>&1
  ncu -i gpurun_out/${R}_prof_$K.ncu-rep --page raw --csv > gpurun_out/${R}_raw_$K.csv 2>/dev/null
  rm -f gpurun_out/${R}_prof_$K.ncu-rep
done
wc -l gpurun_out/${R}_launches.csv gpurun_out/${R}_raw_*.csv | tail -4
