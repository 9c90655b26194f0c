mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn_tc.py tests/test_gpu_edgeconv.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/knn_survivor_counts.py 2>&1 | grep "kNN call" | sed 's/survivors.*flagged/.. flagged/'
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train --skip-kmeans 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches']); print(d['roofline_detail']['entry_point_ms_per_step'])"
