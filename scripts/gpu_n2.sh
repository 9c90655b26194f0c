mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 2500 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n$N.json 2>> gpurun_out/r2_bench_n$N.err
tail -c 600 gpurun_out/r2_bench_reference_n$N.json
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
