#!/bin/bash
# A/B of the in-tree library ("new") against profiles/scripts/libfgd_base.so ("base"): GPU tests with the new one, then the
# four kernels' quick bench lines with both.  Usage: r02_ab.sh TAG [workloads...]
TAG=${1:-r02q}; shift
WL=${@:-c5 c2sat c3 c4}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--no-cpu-baseline --no-e2e --no-parity --no-secondary"
for lib in new base new base; do
  if [ $lib = base ]; then export FGD_LIBRARY=$PWD/profiles/scripts/libfgd_base.so; else unset FGD_LIBRARY; fi
  for w in $WL; do
    case $w in
      c5) A="--workload c5 --batch 262144 --steps 4 --warmup 1";;
      c2sat) A="--workload c2 --batch 65536 --steps 5 --warmup 2";;
      c2) A="--workload c2 --steps 10 --warmup 3";;
      c3) A="--workload c3 --batch 8192 --steps 2 --warmup 1";;
      c4) A="--workload c4 --batch 65536 --steps 2 --warmup 1";;
      c1) A="--workload c1 --steps 20 --warmup 5";;
    esac
    timeout 600 python bench.py $A $B >> gpurun_out/${TAG}_${w}_$lib.json 2>> gpurun_out/${TAG}_${w}_$lib.err
  done
done
unset FGD_LIBRARY
python - $TAG $WL <<'PY'
import json, sys
tag = sys.argv[1]
for w in sys.argv[2:]:
    out = []
    for lib in ("new", "base"):
        try:
            ms = [json.loads(l)["ms_per_step"] for l in open(f"gpurun_out/{tag}_{w}_{lib}.json") if l.strip().startswith("{")]
            out.append((lib, [round(m, 4) for m in ms]))
        except Exception as e:
            out.append((lib, "failed " + str(e)[:80]))
    try:
        print(w, out, "new/base time ratio %.4f" % (min(out[0][1]) / min(out[1][1])))
    except Exception:
        print(w, out)
PY
