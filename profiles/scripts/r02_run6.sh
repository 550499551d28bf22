#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CLK_RUNS="c3:2368,c5:9472,c2:9472" timeout 600 python profiles/scripts/clk.py 0 > gpurun_out/r02f_phase_clocks_loaded.txt 2>&1; cat gpurun_out/r02f_phase_clocks_loaded.txt
B="--no-cpu-baseline --no-parity --no-secondary --presoak-seconds 0.3"
timeout 300 python bench.py --workload c1 --steps 200 --warmup 20 $B > gpurun_out/r02f_c1.json 2> gpurun_out/r02f_c1.err
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02f_c2.json 2> gpurun_out/r02f_c2.err
python - <<'PY'
import json
for n in ("c1","c2"):
    d=json.load(open(f"gpurun_out/r02f_{n}.json")); print(n, d["value"], d["ms_per_step"], d["roofline"]["frac"])
PY
