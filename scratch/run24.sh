#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run24.txt
: > $OUT
timeout 150 python -m pytest tests -m gpu -x -q -k "bit_exact or zero_copy or full_size or device_init or budget or host_buffer" 2>&1 | tail -3 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'it=%.2f'%d['mean_inner_iters'])
"; }
for q in 0 24 32 48; do
  FGD_QUANTUM=$q timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-saturated --presoak-seconds 0.2 2>>gpurun_out/run24.err | summ "quantum=$q c2 B4096" >> $OUT
done
cat $OUT; tail -3 gpurun_out/run24.err
