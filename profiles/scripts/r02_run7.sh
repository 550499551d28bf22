#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="--no-cpu-baseline --no-e2e --no-parity --no-secondary --presoak-seconds 0.3"
for v in 0 1; do
  FGD_VARIANT=$v timeout 600 python bench.py --workload c3 --batch 8192 --steps 2 --warmup 1 $B > gpurun_out/r02g_c3_b8192_v$v.json 2> gpurun_out/r02g_c3_b8192_v$v.err; echo "c3 v$v rc=$?"
done
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 $B > gpurun_out/r02g_c3_full.json 2> gpurun_out/r02g_c3_full.err; echo "c3 full rc=$?"
python - <<'PY'
import json
for n in ("c3_b8192_v0","c3_b8192_v1","c3_full"):
    try:
        d=json.load(open(f"gpurun_out/r02g_{n}.json")); print(n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["launch"])
    except Exception as e: print(n, "failed", e)
PY
