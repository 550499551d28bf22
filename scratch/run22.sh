#!/bin/bash
# round 1h final: tests, default bench line, reference arm, smoke, launch list, ncu capture of the headline batch, full-size runs
mkdir -p gpurun_out
OUT=gpurun_out/run22.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $OUT
timeout 600 python bench.py > gpurun_out/r01h_bench_c2.json 2> gpurun_out/r01h_bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/r01h_bench_c2.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4g frac %.3f e2e %.4g sat %.4g satfrac %.3f cpu %.4g'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['saturated']['value'],d['saturated']['roofline_frac'],d['cpu_baseline']['value']), d['cpu_baseline']['sample'], d['clocks'])" >> $OUT
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01h_bench_reference_arm.json 2>> gpurun_out/r01h_bench_c2.err
python -c "import __graft_entry__ as g; g.smoke()" >> $OUT 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01h_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_launches_c2.log 2>&1
V=0 B=4096 TAG=r01h_c2_b4096 bash scratch/ncu4.sh >> $OUT 2>&1
summ() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'iters/s=%.4g'%d['fgd_iters_per_s'], 'frac=%.3f'%d['roofline']['frac'], 'mean_it=%.1f'%d['mean_inner_iters'], 'ful=%.2f'%d['fulfilled_frac'])
"; }
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run22.err | tee gpurun_out/r01h_bench_c3_full.json | summ "c3 full B65536" >> $OUT
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>>gpurun_out/run22.err | tee gpurun_out/r01h_bench_c4_full.json | summ "c4 full B262144" >> $OUT
timeout 600 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu-baseline 2>>gpurun_out/run22.err | tee gpurun_out/r01h_bench_c5_full.json | summ "c5 full 1M" >> $OUT
timeout 600 python bench.py --workload c1 --steps 20 --warmup 3 2>>gpurun_out/run22.err | tee gpurun_out/r01h_bench_c1.json | summ "c1" >> $OUT
cat $OUT; tail -3 gpurun_out/run22.err gpurun_out/r01h_bench_c2.err
