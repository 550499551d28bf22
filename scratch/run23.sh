#!/bin/bash
# round 1j final: tests, default bench line, launch list, ncu captures of the final T = 50 kernel
mkdir -p gpurun_out
OUT=gpurun_out/run23.txt
: > $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $OUT
timeout 600 python bench.py > gpurun_out/r01j_bench_c2.json 2> gpurun_out/r01j_bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/r01j_bench_c2.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4g frac %.3f e2e %.4g sat %.4g satfrac %.3f cpu %.4g'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['saturated']['value'],d['saturated']['roofline_frac'],d['cpu_baseline']['value']), d['clocks'])" >> $OUT
python -c "import __graft_entry__ as g; g.smoke()" >> $OUT 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated --presoak-seconds 0 > gpurun_out/r01j_launches_c2.log 2>&1
V=0 B=4096 TAG=r01j_c2_b4096 bash scratch/ncu4.sh >> $OUT 2>&1
V=0 B=65536 TAG=r01j_c2_b65536 bash scratch/ncu4.sh >> $OUT 2>&1
cat $OUT
