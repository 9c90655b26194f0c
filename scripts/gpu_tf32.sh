mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --tb=short 2>&1 | grep -v "^model using\|^$\|^     +" | tail -8
timeout 300 python scripts/bench_train.py --steps 10 > gpurun_out/r2j_train.json 2> /dev/null
