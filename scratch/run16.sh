#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run16.txt
: > $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.4g'%(e.get('value') or 0), (e.get('transfer') or '')[:9], d['config'].get('launch'))
"; }
for v in 0 2; do
for b in 4096 65536; do
  FGD_VARIANT=$v timeout 200 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run16.err | summ "v$v c2 B$b" >> $OUT
done
done
cat $OUT; tail -5 gpurun_out/run16.err
