#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "live or dynamic" 2>&1 | tail -3
B="--no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0.5"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 1 $B > gpurun_out/r02k_c4_pipe_tc0.json 2> gpurun_out/r02k_c4_pipe_tc0.err; echo "rc=$?"
FGD_LIVE_FORCE_TC=1 timeout 600 python bench.py --workload c4 --steps 3 --warmup 1 $B > gpurun_out/r02k_c4_block_tc50.json 2> gpurun_out/r02k_c4_block_tc50.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("c4_pipe_tc0","c4_block_tc50"):
    d=json.load(open(f"gpurun_out/r02k_{n}.json")); print(n, d["value"], d["step_ms"], d["roofline"]["frac"], d["dynamic_obstacles"]["sets_published_per_step"])
PY
