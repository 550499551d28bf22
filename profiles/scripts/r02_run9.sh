#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--no-cpu-baseline --no-e2e --no-parity --no-secondary"
for lib in pipe nopipe; do
  if [ $lib = nopipe ]; then export FGD_LIBRARY=$PWD/profiles/scripts/libfgd_nopipe.so; else unset FGD_LIBRARY; fi
  timeout 600 python bench.py --workload c3 --batch 8192 --steps 2 --warmup 1 $B > gpurun_out/r02j_c3_$lib.json 2> gpurun_out/r02j_c3_$lib.err
  timeout 600 python bench.py --workload c4 --batch 65536 --steps 2 --warmup 1 $B > gpurun_out/r02j_c4_$lib.json 2> gpurun_out/r02j_c4_$lib.err
  timeout 600 python bench.py --workload c5 --batch 262144 --steps 3 --warmup 1 $B > gpurun_out/r02j_c5_$lib.json 2> gpurun_out/r02j_c5_$lib.err
  timeout 600 python bench.py --workload c2 --batch 65536 --steps 3 --warmup 1 $B > gpurun_out/r02j_c2sat_$lib.json 2> gpurun_out/r02j_c2sat_$lib.err
done
unset FGD_LIBRARY
python - <<'PY'
import json
for w in ("c3","c4","c5","c2sat"):
    r=[]
    for lib in ("pipe","nopipe"):
        try:
            d=json.load(open(f"gpurun_out/r02j_{w}_{lib}.json")); r.append((round(d["ms_per_step"],3), round(d["roofline"]["frac"],4)))
        except Exception as e: r.append(("failed",str(e)[:60]))
    print(w, "pipe", r[0], "nopipe", r[1])
PY
