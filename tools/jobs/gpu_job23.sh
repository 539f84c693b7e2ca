#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -15 > gpurun_out/j23_tests.log
timeout 300 python tools/attn_bench.py > gpurun_out/j23_attn_bench.txt 2>&1
cat gpurun_out/j23_tests.log gpurun_out/j23_attn_bench.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generate > gpurun_out/j23_bench.json 2> gpurun_out/j23_bench.err
tail -c 3000 gpurun_out/j23_bench.json
