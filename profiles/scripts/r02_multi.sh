#!/bin/bash
# usage: r02_multi.sh N TAG   (N GPUs of one box; torchrun, NCCL)
N=${1:-2}; TAG=${2:-r02e}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
fi
run() {
  NAME=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
    > gpurun_out/${TAG}_${NAME}_n$N.json 2> gpurun_out/${TAG}_${NAME}_n$N.err
  echo "$NAME rc=$?"; grep "\[bench\]" gpurun_out/${TAG}_${NAME}_n$N.err | head -8
}
run c5 --steps 5 --warmup 2
run c2 --workload c2 --steps 10 --warmup 3 --no-parity
[ "$N" = "2" ] && NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 1 --warmup 1 --batch $((N*65536)) --no-e2e --no-parity 2>&1 | grep -E "NVLS|Connected all|via P2P|NET/" | sort | uniq -c | head -12 > gpurun_out/${TAG}_nccl_n$N.txt
python - <<PY
import json
for n in ("c5","c2"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_%s_n$N.json" % n) if l.startswith("{")][-1]); print(n, "N=$N", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("sweep"))
    except Exception as e: print(n, "failed", e)
PY
cat gpurun_out/${TAG}_nccl_n$N.txt
