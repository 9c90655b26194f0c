#!/bin/bash
# Round profile run (one GPU): bench line, smoke, ncu launch list, one --set full capture per hand-written kernel.
# Usage (from the repo root, through gpurun):  bash scripts/gpu_profile.sh r1
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_nvidia_smi.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1
# every launch of two timed steps with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# KERNELS="..." restricts the --set full captures (e.g. to the kernels that changed since the last full run)
KERNELS=${KERNELS:-knn_tc_kernel knn_finish_kernel knn_prep_kernel edgeconv_kernel edge_pq_kernel linear_kernel attention_kernel rowsel_tc_kernel rowsel_recheck_kernel cos_logits_kernel softmax_pool_kernel gemm_tf32_kernel}
for K in $KERNELS; do
  [ $K = gemm_tf32_kernel ] && continue
  case $K in knn_*) SKIP="-s 6 -c 6" ;; *) SKIP="-s 4 -c 1" ;; esac      # the kNN kernels: all three layers (two chains each) of one step (traffic accounting)
  ncu --set full --clock-control none --import-source on -k regex:$K $SKIP -o gpurun_out/${R}_prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  ncu -i gpurun_out/${R}_prof_$K.ncu-rep --page raw --csv > gpurun_out/${R}_raw_$K.csv 2>/dev/null
  # gpurun brings back at most 64 MiB: keep the full report of the two dominant kernels only, the raw pages of all
  case $K in knn_tc_kernel|edgeconv_kernel) ;; *) rm -f gpurun_out/${R}_prof_$K.ncu-rep ;; esac
done
# the training GEMM (tcgen05 kind::tf32, 3xTF32 mode): conv2 forward / data gradient / weight gradient of one EdgeConv layer
case " $KERNELS " in *" gemm_tf32_kernel "*) ;; *) du -sh gpurun_out; exit 0 ;; esac
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 45 -c 10 -o gpurun_out/${R}_prof_gemm_tf32_kernel \
    python scripts/bench_train.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu -i gpurun_out/${R}_prof_gemm_tf32_kernel.ncu-rep --page raw --csv > gpurun_out/${R}_raw_gemm_tf32_kernel.csv 2>/dev/null
rm -f gpurun_out/${R}_prof_gemm_tf32_kernel.ncu-rep
du -sh gpurun_out
python scripts/bench_train.py --steps 10 > gpurun_out/${R}_train_n1.json 2>/dev/null
ls -la gpurun_out | tail -30
