#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep5.txt
: > $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config'].get('launch'))
"; }
for v in ${VARIANTS:-0}; do
export FGD_VARIANT=$v
for b in 4096 65536; do
  timeout 120 python bench.py --workload c2 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep5.err | summ "v$v c2 B$b" >> $OUT
done
timeout 300 python bench.py --workload c3 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep5.err | summ "v$v c3 B4096" >> $OUT
timeout 300 python bench.py --workload c3 --batch 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep5.err | summ "v$v c3 B16384" >> $OUT
done
unset FGD_VARIANT
timeout 300 python bench.py --workload c4 --batch 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep5.err | summ "c4 B65536" >> $OUT
timeout 300 python bench.py --workload c5 --batch 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep5.err | summ "c5 B262144" >> $OUT
cat $OUT
tail -5 gpurun_out/sweep5.err
