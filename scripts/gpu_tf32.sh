mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x 2>&1 | tail -3
timeout 300 python scripts/bench_train.py --steps 10 2>/dev/null > gpurun_out/r2k_train.json; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2k_train.json').read().strip().splitlines()[-1])
e=d['entry_point_ms']
print('step',d['ms_per_step'],'gemm',sum(v for k,v in e.items() if 'gemm' in k), 'mem', d['peak_mem_gb'], 'launches', d['gpu_launches_per_step'])
P
