mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_callers.py 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value']); print(d['roofline_detail']['entry_point_ms_per_step']); print('knn ms', d['roofline']['ms_per_step'])"
timeout 300 python scripts/bench_train.py --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train ms/step', d['ms_per_step'], 'launches', d['gpu_launches_per_step'])"
