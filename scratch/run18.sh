#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run18.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.4g'%(e.get('value') or 0), (e.get('transfer') or '')[:9], d['config'].get('launch'))
"; }
for v in ${VARIANTS:-0 3 4}; do
for b in 4096 65536; do
  FGD_VARIANT=$v timeout 200 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run18.err | summ "v$v c2 B$b" >> $OUT
done
done
cat $OUT; tail -5 gpurun_out/run18.err
