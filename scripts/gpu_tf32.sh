mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_reference_callers.py -q -s -x 2>&1 | tail -30
