#!/bin/bash
# one --set full capture that keeps the source page CSV (per-instruction executed counts and stall samples) in gpurun_out/
TAG=$1; W=$2; B=$3
mkdir -p gpurun_out
ARGS="--workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0"
N=${TAG}_${W}_b${B}
timeout 600 python bench.py $ARGS > gpurun_out/${N}_plain.json 2> gpurun_out/${N}_plain.err || { echo "$N plain run failed"; tail -3 gpurun_out/${N}_plain.err; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgd_optimize_kernel -s 1 -c 1 -o gpurun_out/${N} -f python bench.py $ARGS > gpurun_out/${N}_ncu.log 2>&1
ncu -i gpurun_out/${N}.ncu-rep --page details > gpurun_out/${N}_details.txt 2>&1
ncu -i gpurun_out/${N}.ncu-rep --page source --csv > gpurun_out/${N}_source.csv 2>&1
ncu -i gpurun_out/${N}.ncu-rep --page raw --csv > gpurun_out/${N}_raw.csv 2>&1
python profiles/scripts/sass_hist.py gpurun_out/${N}_source.csv > gpurun_out/${N}_sass_histogram.txt 2>&1
rm -f gpurun_out/${N}.ncu-rep
echo "$N done"
