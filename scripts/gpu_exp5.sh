mkdir -p gpurun_out
GFS3D_KNN_ATMEM=1 timeout 300 python -m pytest tests/test_gpu_knn_tc.py -m gpu -q -x 2>&1 | tail -3
for a in 0 1; do
GFS3D_KNN_ATMEM=$a timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-train --skip-kmeans 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('atmem $a blocks/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'knn', round(d['roofline_detail']['entry_point_ms_per_step']['gfs_knn_tc_set_f32'],4))"
done
