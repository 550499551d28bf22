#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/run12.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $OUT
timeout 600 python bench.py > gpurun_out/r01g_bench_c2.json 2> gpurun_out/r01g_bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/r01g_bench_c2.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4g frac %.3f e2e %.4g sat %.4g satfrac %.3f cpu %.4g'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['saturated']['value'],d['saturated']['roofline_frac'],d['cpu_baseline']['value']), d['cpu_baseline']['sample'])" >> $OUT
python -c "import __graft_entry__ as g; g.smoke()" >> $OUT 2>&1
cat $OUT; tail -3 gpurun_out/r01g_bench_c2.err
