#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run9.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config']['launch'])
"; }
timeout 120 python bench.py --workload c2 --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run9.err | summ "c2 B65536" >> $OUT
timeout 120 python bench.py --workload c2 --batch 65536 --whole-arm --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run9.err | summ "c2 B65536 whole-arm" >> $OUT
timeout 300 python bench.py --workload c5 --batch 131072 --whole-arm --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run9.err | summ "c5 B131072 whole-arm" >> $OUT
timeout 300 python bench.py --workload c3 --batch 4096 --whole-arm --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run9.err | summ "c3 B4096 whole-arm" >> $OUT
timeout 300 python main.py --whole-arm-cost true 2>&1 | tail -6 >> $OUT
cat $OUT; tail -3 gpurun_out/run9.err
