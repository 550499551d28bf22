#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run6.txt
: > $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
timeout 600 python bench.py > gpurun_out/r01e_bench_c2.json 2> gpurun_out/r01e_bench_c2.err
cat gpurun_out/r01e_bench_c2.json >> $OUT
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01e_bench_reference_arm.json 2>> gpurun_out/r01e_bench_c2.err
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], d['config'].get('launch'))
"; }
timeout 300 python bench.py --workload c3 --batch 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run6.err | tee gpurun_out/r01e_bench_c3_b16384.json | summ "c3 B16384" >> $OUT
timeout 300 python bench.py --workload c4 --batch 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run6.err | summ "c4 B65536" >> $OUT
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/run6.err | summ "c1" >> $OUT
cat $OUT | cut -c1-1800
tail -3 gpurun_out/run6.err gpurun_out/r01e_bench_c2.err
