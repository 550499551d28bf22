#!/bin/bash
# ncu --set full of the T = 50 optimiser kernel, tables in tensor memory (variant $V) - c2 shape, B = 65536
mkdir -p gpurun_out
run() {
  V=$1; W=$2; B=$3; TAG=$4
  FGD_VARIANT=$V timeout 600 python bench.py --workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --presoak-seconds 0 > gpurun_out/${TAG}_plain.log 2>&1 || return 1
  FGD_VARIANT=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgd_optimize_kernel -s 1 -c 1 -o gpurun_out/${TAG} -f python bench.py --workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --presoak-seconds 0 > gpurun_out/${TAG}_ncu.log 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>&1
  rm -f gpurun_out/${TAG}.ncu-rep
  tail -2 gpurun_out/${TAG}_ncu.log
}
run ${V:-0} c2 ${B:-65536} ${TAG:-r01h_c2_b65536}
