mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2_kmeans_launches.csv python scripts/bench_kmeans.py --n 500000 --iters 4 --cpu-sample 0 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_kmeans_launches.csv')) if len(r)>10]
h=rows[0]; iN=h.index("Kernel Name"); iV=h.index("Metric Value")
for r in rows[1:45]: print(f"{float(r[iV])/1000:8.1f} {r[iN][:90]}")
P
