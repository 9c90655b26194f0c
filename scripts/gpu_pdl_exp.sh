# programmatic dependent launch: the same short bench line with the attribute off / single-stream kernels only / all
# kernels (GFS3D_PDL = 0, 1, 3), then the smoke test with it on.
mkdir -p gpurun_out
for P in 0 1 3 0 3; do
GFS3D_PDL=$P timeout 120 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --skip-train --skip-kmeans 2>gpurun_out/pdl_$P.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL=$P blocks/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/pdl_$P.err
done
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
