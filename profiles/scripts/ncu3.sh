#!/bin/bash
# ncu --set full captures (second launch = first timed step) + launch list of the default bench
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
run c2 65536 r01d_c2_b65536
run c2 4096 r01d_c2_b4096
run c3 4096 r01d_c3_b4096
python bench.py --steps 3 --warmup 3 > gpurun_out/r01d_bench_c2.json 2> gpurun_out/r01d_bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_launches_c2.log 2>&1
tail -3 gpurun_out/r01d_bench_c2.json | cut -c1-600
