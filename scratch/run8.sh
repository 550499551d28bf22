#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run8.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config']['launch'])
"; }
for v in 0 2 3; do
  export FGD_VARIANT=$v
  for b in 4096 65536; do
    timeout 120 python bench.py --workload c2 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run8.err | summ "v$v c2 B$b" >> $OUT
  done
  timeout 300 python bench.py --workload c5 --batch 131072 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run8.err | summ "v$v c5 B131072" >> $OUT
done
unset FGD_VARIANT
timeout 300 python main.py --replan 20 2>&1 | tail -4 >> $OUT
timeout 300 python main.py --replan 10 --batch 4096 2>&1 | tail -3 >> $OUT
cat $OUT; tail -3 gpurun_out/run8.err
