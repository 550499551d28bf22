mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02zf_bench_default.json 2> gpurun_out/r02zf_bench_default.err; echo "bench rc=$?"
bash profiles/scripts/r02_counters.sh r02zf c4:32768:full
