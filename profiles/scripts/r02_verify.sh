mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02zj_pytest.txt
echo "pytest took $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02zj_bench_default.json 2> gpurun_out/r02zj_bench_default.err
echo "bench rc=$? took $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02zj_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
