#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run10.txt
: > $OUT
FGD_VARIANT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config']['launch'])
"; }
for v in 0 2; do
  export FGD_VARIANT=$v
  for b in 1 4096 8192 65536; do
    timeout 120 python bench.py --workload c2 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run10.err | summ "v$v c2 B$b" >> $OUT
  done
  timeout 300 python bench.py --workload c5 --batch 131072 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run10.err | summ "v$v c5 B131072" >> $OUT
  timeout 300 python bench.py --workload c4 --batch 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run10.err | summ "v$v c4 B65536" >> $OUT
  timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run10.err | summ "v$v c1" >> $OUT
done
cat $OUT; tail -3 gpurun_out/run10.err
