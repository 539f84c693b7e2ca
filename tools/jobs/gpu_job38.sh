#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
timeout 300 python tools/attn_bench.py > gpurun_out/j38_attn_bench.txt 2>&1
cat gpurun_out/j38_attn_bench.txt
