# ncu source-level hot spots of one kernel of the inference step:  bash scripts/gpu_src_profile.sh <kernel regex> [skip]
K=${1:-rowsel_tc_kernel}
S=${2:-2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/src_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-train --skip-kmeans > /dev/null 2>&1
ncu -i gpurun_out/src_$K.ncu-rep --page source --csv > gpurun_out/src_$K.csv 2>/dev/null
rm -f gpurun_out/src_$K.ncu-rep
wc -l gpurun_out/src_$K.csv
