#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02a_pytest.txt
tail -5 gpurun_out/r02a_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 2 > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r02a_bench_default.err; cut -c1-1500 gpurun_out/r02a_bench_default.json
bash profiles/scripts/r02_counters.sh r02a c5:65536:full c2:65536 c2:4096 c3:4096:full c4:32768
