#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run21.txt
: > $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'it=%.2f'%d['mean_inner_iters'])
"; }
for lib in irm_motion_planning_b200/libfgd_b200.so scratch/libfgd_ld32.so; do
for b in 4096 65536; do
FGD_LIBRARY=$PWD/$lib timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run21.err | summ "$lib c2 B$b" >> $OUT
done
FGD_LIBRARY=$PWD/$lib timeout 300 python bench.py --workload c5 --batch 262144 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run21.err | summ "$lib c5 B262144" >> $OUT
done
cat $OUT; tail -5 gpurun_out/run21.err
