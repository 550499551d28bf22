#!/bin/bash
# torchrun on N GPUs: c2 weak scaling (default bench line) and the 1 M restart sweep (c5)
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/multi_h_$N.txt
: > $OUT
nvidia-smi -L >> $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/multi_h_$N.err | grep '^{' | tee gpurun_out/r01h_scale_c2_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('c2 N=%d'%d['n_gpus'], 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.4g'%d['e2e']['value'], d['e2e']['transfer'][:9])" >> $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 3 --warmup 2 --no-e2e 2>>gpurun_out/multi_h_$N.err | grep '^{' | tee gpurun_out/r01h_scale_c5_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('c5 1M strong N=%d'%d['n_gpus'], 'value=%.4g'%d['value'], 'ms=%.4g'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'])" >> $OUT
cat $OUT; tail -5 gpurun_out/multi_h_$N.err
