#!/bin/bash
mkdir -p gpurun_out
ATTN_ONE=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tiled -s 3 -c 3 -o gpurun_out/j22_attn_tiled -f python tools/attn_bench.py > gpurun_out/j22_ncu.log 2>&1
tail -3 gpurun_out/j22_ncu.log
