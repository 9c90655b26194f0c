mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --tb=short 2>&1 | grep -v "^model using\|^$\|^     +" | tail -30
timeout 300 python scripts/bench_train.py --steps 5 > gpurun_out/r2i_train.json 2> gpurun_out/r2i_train.err; tail -3 gpurun_out/r2i_train.err
