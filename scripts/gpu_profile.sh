mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 3000 gpurun_out/bench_r1_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches_r1.csv
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 7 -c 2 -o gpurun_out/prof_knn_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_knn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edgeconv_kernel -s 4 -c 1 -o gpurun_out/prof_edgeconv_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 6 -c 2 -o gpurun_out/prof_linear_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lin.log 2>&1
ls -la gpurun_out
