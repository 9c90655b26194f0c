# what the driver runs at round end, in one call: the GPU suite (reference callers included), smoke, both bench arms
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reference', d['value'], d['cpu_baseline']['kind'])"
python bench.py 2>/dev/null > gpurun_out/r2_bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print('bench', d['value'], d['e2e']['value'], d['train']['ms_per_step'], d['kmeans']['ms_per_iter'], d['roofline_detail']['knn_graph_tensor'])"
