#!/bin/bash
mkdir -p gpurun_out
ATTN_ONE=1 timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tiled -s 2 -c 2 -o gpurun_out/j49_attn_vit -f python tools/attn_bench.py > gpurun_out/j49_ncu_vit.log 2>&1
ATTN_ONE=1 ATTN_T=64 ATTN_H=16 ATTN_CAUSAL=1 timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tiled -s 2 -c 2 -o gpurun_out/j49_attn_mbart -f python tools/attn_bench.py > gpurun_out/j49_ncu_mbart.log 2>&1
tail -2 gpurun_out/j49_ncu_vit.log gpurun_out/j49_ncu_mbart.log
