mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kmeans.py tests/test_gpu_fullsize.py tests/test_gpu_callers.py -q -s --tb=short 2>&1 | grep -v "^model using\|^$" | tail -25
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value']); print(d['roofline_detail']['entry_point_ms_per_step']); print('knn ms', d['roofline']['ms_per_step'])"
