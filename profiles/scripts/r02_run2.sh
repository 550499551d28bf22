#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "live or config5 or two_streams or main_report or per_trajectory or convergence" 2>&1 | tail -30 > gpurun_out/r02b_pytest_new.txt
tail -12 gpurun_out/r02b_pytest_new.txt
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0.5"
timeout 600 python bench.py --workload c4 $B > gpurun_out/r02b_c4_live.json 2> gpurun_out/r02b_c4_live.err; echo "c4 live rc=$?"; tail -2 gpurun_out/r02b_c4_live.err
timeout 600 python bench.py --workload c4 --c4-relaunch $B > gpurun_out/r02b_c4_relaunch.json 2> gpurun_out/r02b_c4_relaunch.err; echo "c4 relaunch rc=$?"
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02b_c2.json 2> gpurun_out/r02b_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c2 --batch 65536 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02b_c2sat.json 2> gpurun_out/r02b_c2sat.err; echo "c2sat rc=$?"
python - <<'PY'
import json
for n in ("c4_live","c4_relaunch","c2","c2sat"):
    try:
        d=json.load(open(f"gpurun_out/r02b_{n}.json")); print(n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("dynamic_obstacles"))
    except Exception as e: print(n, "failed", e)
PY
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
