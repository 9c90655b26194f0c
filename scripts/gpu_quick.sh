# Quick GPU check during development (through gpurun): the parity suites that cover the kernels being edited, a short
# inference bench line, the smoke test and the training bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_callers.py 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train --skip-kmeans 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches']); print(d['roofline_detail']['entry_point_ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/bench_train.py --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train ms/step', d['ms_per_step'], 'peak GB', d['peak_mem_gb'], 'launches', d['gpu_launches_per_step'])"
