#!/bin/bash
# A/B/n of library builds: r02_abn.sh TAG "lib1 lib2 ..." "workloads" ; libX = name of profiles/scripts/libfgd_<name>.so, "tree" = the in-tree build
TAG=$1; LIBS=$2; WL=${3:-"c5 c2sat"}
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-parity --no-secondary"
for rep in 1 2; do
for lib in $LIBS; do
  if [ $lib = tree ]; then unset FGD_LIBRARY; else export FGD_LIBRARY=$PWD/profiles/scripts/libfgd_$lib.so; fi
  for w in $WL; do
    case $w in
      c5) A="--workload c5 --batch 262144 --steps 4 --warmup 1";;
      c2sat) A="--workload c2 --batch 65536 --steps 5 --warmup 2";;
      c2) A="--workload c2 --steps 10 --warmup 3";;
      c3) A="--workload c3 --batch 8192 --steps 2 --warmup 1";;
      c4) A="--workload c4 --batch 65536 --steps 2 --warmup 1";;
      c1) A="--workload c1 --steps 20 --warmup 5";;
    esac
    timeout 600 python bench.py $A $B $EXTRA >> gpurun_out/${TAG}_${w}_$lib.json 2>> gpurun_out/${TAG}_${w}_$lib.err
  done
done
done
unset FGD_LIBRARY
python - $TAG "$LIBS" "$WL" <<'PY'
import json, sys
tag, libs, wls = sys.argv[1], sys.argv[2].split(), sys.argv[3].split()
for w in wls:
    base = None
    for lib in libs:
        try:
            ms = [json.loads(l)["ms_per_step"] for l in open(f"gpurun_out/{tag}_{w}_{lib}.json") if l.strip().startswith("{")]
            if base is None: base = min(ms)
            print(w, lib, [round(m, 4) for m in ms], "vs first %.4f" % (min(ms) / base))
        except Exception as e:
            print(w, lib, "failed", str(e)[:80])
PY
