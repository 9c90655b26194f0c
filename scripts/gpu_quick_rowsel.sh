mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kmeans.py tests/test_gpu_fullsize.py tests/test_gpu_edgeconv.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value']); print(d['roofline_detail']['entry_point_ms_per_step']); print(d['kmeans']['ms_per_iter'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
