#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/scripts/t128_ab.py > gpurun_out/r02o_t128_ab.txt 2>&1; cat gpurun_out/r02o_t128_ab.txt | tail -4
timeout 300 python profiles/scripts/sanitize.py > gpurun_out/r02o_all_kernels_smoke.txt 2>&1; tail -2 gpurun_out/r02o_all_kernels_smoke.txt
