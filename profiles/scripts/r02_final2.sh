#!/bin/bash
# Final evidence after the default-scene instance and the unrolled pipelined loop (round 2zo): tests, bench line, smoke, launch list,
# --set full captures of the c5 / c2 / c3 kernels (the c4 capture of r02zg stays valid: that kernel did not change).
TAG=${1:-r02zo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_default.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/${TAG}_launches_default.log 2>&1
echo "launch list rc=$?"
bash profiles/scripts/r02_counters.sh ${TAG} c5:65536:full c2:65536:full c3:4096:full
