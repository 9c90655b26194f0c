#!/bin/bash
# Round profile run (one GPU): bench line, smoke, ncu launch list, one --set full capture per hand-written kernel.
# Usage (from the repo root, through gpurun):  bash scripts/gpu_profile.sh r1
R=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_nvidia_smi.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1
# every launch of two timed steps with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for K in knn_tc_kernel knn_finish_kernel knn_prep_kernel edgeconv_kernel linear_kernel attention_kernel rowsel_kernel pointwise_kernel cos_logits_kernel softmax_pool_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -o gpurun_out/${R}_prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  ncu -i gpurun_out/${R}_prof_$K.ncu-rep --page raw --csv > gpurun_out/${R}_raw_$K.csv 2>/dev/null
done
ls -la gpurun_out | tail -30
