#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py > gpurun_out/j19_attn_bench.txt 2>&1
ATTN_ONE=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tiled -s 3 -c 3 -o gpurun_out/j19_attn_tiled -f python tools/attn_bench.py > gpurun_out/j19_ncu.log 2>&1
tail -3 gpurun_out/j19_ncu.log
cat gpurun_out/j19_attn_bench.txt
