#!/bin/bash
# attention A/B: parity tests of both kernel families, then the timing table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -15 > gpurun_out/j18_tests.log
timeout 300 python tools/attn_bench.py > gpurun_out/j18_attn_bench.txt 2>&1
cat gpurun_out/j18_tests.log gpurun_out/j18_attn_bench.txt
