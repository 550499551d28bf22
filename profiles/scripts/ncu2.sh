#!/bin/bash
# ncu --set full captures of the optimiser kernel on the obstacle-dominated shapes (c3, c4)
mkdir -p gpurun_out
run() {
  W=$1; B=$2; TAG=$3
  timeout 600 python bench.py --workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_plain.log 2>&1 || return 1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgd_optimize_kernel -s 1 -c 1 -o gpurun_out/${TAG} -f python bench.py --workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu.log 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>&1
  rm -f gpurun_out/${TAG}.ncu-rep
  tail -2 gpurun_out/${TAG}_ncu.log
}
run c3 2048 r01c_c3_b2048
run c4 32768 r01c_c4_b32768
