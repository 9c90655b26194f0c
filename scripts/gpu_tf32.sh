mkdir -p gpurun_out
for CFG in "0 0" "1 0" "1 1"; do
set -- $CFG
GFS_OVERLAP=$1 GFS_KNN_SPLIT=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train --skip-kmeans > gpurun_out/ov.json 2> gpurun_out/ov.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/ov.json').read()); print('overlap $1 split $2 blocks/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/ov.err
done
GFS_KNN_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_edgeconv.py -q -x 2>&1 | tail -3
