#!/bin/bash
# Final single-GPU evidence of a round: tests, the default bench line, its ncu launch list, --set full captures.
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
echo "reference arm rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# launch list of the same command (shorter run, primary workload only: with the secondary workloads the serialised
# 4 s kernels of c3 and the live publisher of c4 ran into the 25-minute limit in round 2h - the CSV of that run covers
# the config-5 part and the peak probes; a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_default.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/${TAG}_launches_default.log 2>&1
echo "launch list rc=$?"
bash profiles/scripts/r02_counters.sh ${TAG} c3:4096:full c5:65536:full c2:65536:full c4:32768:full
