#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run19.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $OUT
FGD_VARIANT=1 timeout 600 python -m pytest tests -m gpu -x -q -k "bit_exact or zero_copy or resume or dynamic" 2>&1 | tail -3 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'it=%.2f'%d['mean_inner_iters'], 'e2e=%.4g'%(e.get('value') or 0), (e.get('transfer') or '')[:9], d['config'].get('launch'))
"; }
for b in 4096 65536; do
  timeout 200 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/run19.err | summ "c2 B$b" >> $OUT
done
timeout 300 python bench.py --workload c5 --batch 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run19.err | summ "c5 B262144" >> $OUT
timeout 300 python bench.py --workload c4 --batch 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run19.err | summ "c4 B65536" >> $OUT
timeout 300 python bench.py --workload c3 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run19.err | summ "c3 B4096" >> $OUT
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/run19.err | summ "c1" >> $OUT
cat $OUT; tail -5 gpurun_out/run19.err
