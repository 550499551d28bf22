#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run14.txt
: > $OUT
FGD_VARIANT=0 timeout 200 python scratch/latency.py >> $OUT 2>&1
FGD_VARIANT=1 timeout 200 python scratch/latency.py >> $OUT 2>&1
cat $OUT
