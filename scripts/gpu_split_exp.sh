# kNN two-chain experiment: the new parity test, then the inference bench line under several splits
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn_tc.py -m gpu -q -x 2>&1 | tail -3
for s in 0 -1 18 12 20; do
  GFS3D_KNN_SPLIT=$s timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train --skip-kmeans 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('split $s blocks/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'knn', round(d['roofline_detail']['entry_point_ms_per_step']['gfs_knn_tc_set_f32'],4))"
done
