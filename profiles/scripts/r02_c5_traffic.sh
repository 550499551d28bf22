#!/bin/bash
# DRAM bytes of the FULL-SIZE config-5 launch (1 M trajectories): a counters-only ncu pass (two metrics, one replay),
# after the plain command exited 0.  A number printed under ncu is never a bench value.
TAG=${1:-r02z}
ARGS="--workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0"
timeout 600 python bench.py $ARGS > gpurun_out/${TAG}_c5_full_plain.json 2> gpurun_out/${TAG}_c5_full_plain.err || { echo plain run failed; exit 1; }
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fgd_optimize_kernel -s 1 -c 1 --csv \
  --log-file gpurun_out/${TAG}_c5_full_dram.csv python bench.py $ARGS > gpurun_out/${TAG}_c5_full_dram.log 2>&1
tail -4 gpurun_out/${TAG}_c5_full_dram.csv
