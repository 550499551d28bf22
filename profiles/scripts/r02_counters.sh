#!/bin/bash
# Executed-FP32 / MUFU instruction counters and --set full captures of the current kernels, one workload at a time.
# usage: r02_counters.sh TAG  (writes gpurun_out/${TAG}_*).  Each capture runs only after the plain command exited 0.
TAG=${1:-r02a}
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum,sm__cycles_elapsed.max,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
run() {
  W=$1; B=$2; FULL=$3
  ARGS="--workload $W --batch $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0"
  # config 4 (live obstacle updates): a launch without updates, so that every replay pass of ncu sees the same work
  [ "$W" = "c4" ] && ARGS="$ARGS --c4-max-sets 0"
  N=${TAG}_${W}_b${B}
  timeout 600 python bench.py $ARGS > gpurun_out/${N}_plain.json 2> gpurun_out/${N}_plain.err || { echo "$N plain run failed"; tail -3 gpurun_out/${N}_plain.err; return 1; }
  timeout 900 ncu --metrics $M --clock-control none -k regex:fgd_optimize_kernel -s 1 -c 1 --csv --log-file gpurun_out/${N}_counters.csv python bench.py $ARGS > gpurun_out/${N}_counters.json 2> gpurun_out/${N}_counters.err
  if [ "$FULL" = "full" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgd_optimize_kernel -s 1 -c 1 -o gpurun_out/${N} -f python bench.py $ARGS > gpurun_out/${N}_ncu.log 2>&1
    ncu -i gpurun_out/${N}.ncu-rep --page details > gpurun_out/${N}_details.txt 2>&1
    ncu -i gpurun_out/${N}.ncu-rep --page source --csv > gpurun_out/${N}_source.csv 2>&1
    ncu -i gpurun_out/${N}.ncu-rep --page raw --csv > gpurun_out/${N}_raw.csv 2>&1
    python profiles/scripts/sass_hist.py gpurun_out/${N}_source.csv > gpurun_out/${N}_sass_histogram.txt 2>&1
    rm -f gpurun_out/${N}.ncu-rep gpurun_out/${N}_source.csv
  fi
  echo "$N done"
}
shift
for spec in "$@"; do
  IFS=: read W B FULL <<< "$spec"
  run $W $B $FULL
done
