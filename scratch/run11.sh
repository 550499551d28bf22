#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run11.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], 'ful=%.2f'%d['fulfilled_frac'])
"; }
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run11.err | tee gpurun_out/r01f_bench_c3_full.json | summ "c3 full B65536" >> $OUT
timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run11.err | tee gpurun_out/r01f_bench_c4_full.json | summ "c4 full B262144" >> $OUT
timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run11.err | tee gpurun_out/r01f_bench_c5_full.json | summ "c5 full 1M" >> $OUT
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 2>>gpurun_out/run11.err | tee gpurun_out/r01f_bench_c1.json | summ "c1" >> $OUT
cat $OUT; tail -3 gpurun_out/run11.err
