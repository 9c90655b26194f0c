mkdir -p gpurun_out
timeout 300 python scripts/knn_survivor_counts.py 2>&1 | grep "kNN call"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_exp.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --skip-train --skip-kmeans > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_exp.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[h]; kn=H.index('Kernel Name'); mv=H.index('Metric Value')
out=[(r[kn][:48], float(r[mv].replace(',',''))/1e3) for r in rows[h+1:] if len(r)>mv]
for o in out[-45:]: print(f"{o[0]:50s} {o[1]:8.1f}")
PY
