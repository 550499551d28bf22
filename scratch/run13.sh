#!/bin/bash
# round 1g: staged / zero-copy trajectory I/O, 20- and 24-warp variants
mkdir -p gpurun_out
OUT=gpurun_out/run13.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.4g'%(e.get('value') or 0), (e.get('transfer') or '')[:9], d['config'].get('launch'))
"; }
for b in 4096 65536; do
  FGD_HOST_IO=copy timeout 200 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/run13.err | summ "copy B$b" >> $OUT
  timeout 200 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/run13.err | summ "zc   B$b" >> $OUT
done
for v in 2 3; do
for b in 4096 65536; do
  FGD_VARIANT=$v timeout 200 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run13.err | summ "v$v B$b" >> $OUT
done
done
timeout 300 python bench.py --workload c5 --batch 262144 --steps 2 --warmup 1 --no-cpu-baseline 2>>gpurun_out/run13.err | summ "c5 B262144" >> $OUT
cat $OUT; tail -5 gpurun_out/run13.err
