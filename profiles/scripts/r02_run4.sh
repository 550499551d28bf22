#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/scripts/clk.py 0 592 > gpurun_out/r02d_phase_clocks.txt 2>&1; cat gpurun_out/r02d_phase_clocks.txt
bash profiles/scripts/r02_counters.sh r02d c4:32768:full
