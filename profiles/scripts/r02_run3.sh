#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "speculative or live" 2>&1 | tail -30 > gpurun_out/r02c_pytest_new.txt
tail -8 gpurun_out/r02c_pytest_new.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="--no-cpu-baseline --no-parity --no-secondary --presoak-seconds 0.3"
for spec in 592 0; do
  FGD_SPEC_MAX_BATCH=$spec timeout 300 python bench.py --workload c1 --steps 200 --warmup 20 $B > gpurun_out/r02c_c1_spec$spec.json 2> gpurun_out/r02c_c1_spec$spec.err; echo "c1 spec=$spec rc=$?"
  for bb in 12 148 296 592; do
    FGD_SPEC_MAX_BATCH=$spec timeout 300 python bench.py --workload c5 --batch $((bb*256)) --steps 1 --warmup 0 --no-e2e $B > /dev/null 2>&1
  done
done
python - <<'PY'
import json
for n in ("c1_spec592","c1_spec0"):
    try:
        d=json.load(open(f"gpurun_out/r02c_{n}.json")); print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["launch"], d["mean_inner_iters"])
    except Exception as e: print(n, "failed", e)
PY
# latency vs batch, spec vs sequential (BLS, default scene, random start/goal)
timeout 600 python profiles/scripts/spec_latency.py > gpurun_out/r02c_spec_latency.txt 2>&1; cat gpurun_out/r02c_spec_latency.txt
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0.5"
timeout 600 python bench.py --workload c4 $B > gpurun_out/r02c_c4_live.json 2> gpurun_out/r02c_c4_live.err; echo "c4 live rc=$?"; tail -2 gpurun_out/r02c_c4_live.err
timeout 600 python bench.py --workload c4 --c4-relaunch $B > gpurun_out/r02c_c4_relaunch.json 2> gpurun_out/r02c_c4_relaunch.err; echo "c4 relaunch rc=$?"
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02c_c2.json 2> gpurun_out/r02c_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c2 --batch 65536 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02c_c2sat.json 2> gpurun_out/r02c_c2sat.err; echo "c2sat rc=$?"
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-secondary > gpurun_out/r02c_c5.json 2> gpurun_out/r02c_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
for n in ("c4_live","c4_relaunch","c2","c2sat","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02c_{n}.json")); print(n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("dynamic_obstacles"))
    except Exception as e: print(n, "failed", e)
PY
