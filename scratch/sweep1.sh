#!/bin/bash
# variant sweep + FFMA2 probe (scratch; results -> gpurun_out/sweep1.txt)
mkdir -p gpurun_out
OUT=gpurun_out/sweep1.txt
: > $OUT
./scratch/probe_ffma2 >> $OUT 2>&1
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config'].get('launch'))
"; }
for v in 0 1 3 4 5 6; do
  for b in 4096 65536; do
    FGD_VARIANT=$v timeout 120 python bench.py --workload c2 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c2 v$v B$b" >> $OUT
  done
done
FGD_VARIANT=0 timeout 120 python bench.py --workload c2 --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c2 v0 B1" >> $OUT
FGD_VARIANT=0 timeout 120 python bench.py --workload c2 --batch 592 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c2 v0 B592" >> $OUT
for v in 0 1 2; do
  FGD_VARIANT=$v timeout 300 python bench.py --workload c3 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c3 v$v B4096" >> $OUT
done
timeout 300 python bench.py --workload c4 --batch 65536 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c4 B65536" >> $OUT
timeout 300 python bench.py --workload c5 --batch 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c5 B262144" >> $OUT
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/sweep1.err | summ "c1" >> $OUT
cat $OUT
