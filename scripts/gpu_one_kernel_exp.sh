# one-kernel experiment: short bench line twice (per-entry-point times included), smoke test, the model parity suite
mkdir -p gpurun_out
for P in 1 2; do
timeout 120 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --skip-train --skip-kmeans 2>gpurun_out/exp6_$P.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run $P blocks/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1)); print({k: round(v,4) for k,v in d['roofline_detail']['entry_point_ms_per_step'].items()})" || tail -5 gpurun_out/exp6_$P.err
done
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
