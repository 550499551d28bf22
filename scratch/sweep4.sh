#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep4.txt
: > $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config'].get('launch'))
"; }
for v in 0 3 4 5; do
  export FGD_VARIANT=$v
  for b in 4096 65536; do
    timeout 120 python bench.py --workload c2 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep4.err | summ "v$v c2 B$b" >> $OUT
  done
  timeout 300 python bench.py --workload c5 --batch 131072 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep4.err | summ "v$v c5 B131072" >> $OUT
done
FGD_VARIANT=4 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 >> $OUT
cat $OUT
