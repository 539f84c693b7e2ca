#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py > gpurun_out/j20_attn_bench.txt 2>&1
ATTN_EAGER=1 timeout 900 ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --clock-control none --kernel-name-base demangled -k regex:attention --csv --page raw --log-file gpurun_out/j20_attn_ncu.csv python tools/attn_bench.py > gpurun_out/j20_ncu.log 2>&1
tail -3 gpurun_out/j20_ncu.log
cat gpurun_out/j20_attn_bench.txt
